"""Fraction of the k-means labels that change in each iteration at the benchmark shape (the numbers
behind the incremental M-step):  python tools/label_change_rate.py"""
import sys, torch
sys.path.insert(0, '.')
from hsg_b200 import ops
from hsg_b200.utils.segsort import common as S
torch.manual_seed(235)
dev = torch.device('cuda:0')
for dist in ('iid', 'planted'):
    b, d, s = 6, 256, 448
    if dist == 'iid':
        emb = torch.randn(b, d, s, s, device=dev)
    else:
        emb = torch.empty(b, d, s, s, device=dev)
        blk = (s + 7) // 8
        for i in range(b):
            c = torch.randn(64, d, device=dev); c = c / c.norm(dim=1, keepdim=True)
            yy = torch.arange(s, device=dev) // blk
            idx = (yy.view(-1, 1) * 8 + yy.view(1, -1)).reshape(-1)
            e = c[idx] + 0.5 * torch.randn(s * s, d, device=dev) / d ** 0.5
            emb[i] = e.t().reshape(d, s, s)
    ex = S.segment_by_kmeans_ex(emb, None, [16, 16], iterations=0)
    x = ex['embeddings_with_loc']
    init = S._grid_init([16, 16], (s, s), dev)[0].repeat(b)
    xh, xerr = ops.make_half_copy(x, 256)
    prev = init
    out = []
    for t in range(1, 11):
        lab = ops.kmeans(x, init, 256, t, seg_offsets=ex['seg_offsets'], max_seg_len=s * s, xh=xh, xerr=xerr)
        out.append(float((lab != prev).float().mean()))
        prev = lab
    print(dist, 'fraction of labels changed per iteration:', ' '.join('%.3f' % v for v in out))
