import sys, time, torch
sys.path.insert(0, '.')
from hsg_b200 import ops, _lib
from hsg_b200.utils.segsort import common as S
torch.manual_seed(235)
dev = torch.device('cuda:0')
emb = torch.randn(8, 256, 448, 448, device=dev)
for iters in (1, 3, 6):
    ex = S.segment_by_kmeans_ex(emb, None, [16, 16], iterations=0)
    x = ex['embeddings_with_loc']
    init = S._grid_init([16, 16], (448, 448), dev)[0].repeat(8)
    xh, xerr = ops.make_half_copy(x, 256)
    lab, cent = ops.kmeans(x, init, 256, iters, seg_offsets=ex['seg_offsets'], max_seg_len=448*448, xh=xh, xerr=xerr, return_centroids=True)
    for flags, name in ((_lib.KMEANS_FORCE_TC, 'tc'), (_lib.KMEANS_FORCE_SIMT, 'simt')):
        torch.cuda.synchronize(); t0 = time.time()
        l2, nre = ops.kmeans_estep(x, cent, seg_offsets=ex['seg_offsets'], max_seg_len=448*448, xh=xh, xerr=xerr, flags=flags, return_rechecked=True)
        torch.cuda.synchronize(); dt = time.time() - t0
        print(iters, name, 'N', x.shape[0], 'rechecked', nre.tolist(), 'frac', float(nre[0]) / x.shape[0], 'ms', dt * 1e3, 'xerr mean', float(xerr.mean()))
    sims = (x[:4096] @ cent[0].t())
    top = sims.topk(3, dim=1).values
    print('  gap12 median', float((top[:,0]-top[:,1]).median()), 'P(gap<7e-4)', float(((top[:,0]-top[:,1])<7e-4).float().mean()), 'P(gap13<7e-4)', float(((top[:,0]-top[:,2])<7e-4).float().mean()))
