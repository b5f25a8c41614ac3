"""A/B check of the delta gather variants (run once per variant: the switch is read once per process).
Prints a checksum of the cluster ids of segment_by_kmeans at a config-2-like shape and the time per call;
the sliced and the whole-row gather add in the same order, so the checksums must be equal."""
import sys, time
import torch
sys.path.insert(0, '.')
from hsg_b200.utils.segsort import common as S

b, d, hw, grid, iters = (int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (8, 256, 448, 16, 8)))
torch.manual_seed(3)
emb = torch.randn(b, d, hw, hw, device='cuda')
res = S.segment_by_kmeans(emb, None, [grid, grid], iterations=iters)
ids = res[3]
w = torch.arange(ids.numel(), device='cuda', dtype=torch.int64) % 1000003 + 1
print('checksum', int((ids.to(torch.int64) * w).sum()), 'distinct', int(ids.max()) + 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
  S.segment_by_kmeans(emb, None, [grid, grid], iterations=iters)
torch.cuda.synchronize()
print('ms per call %.2f' % ((time.perf_counter() - t0) / 3 * 1e3))
