"""Per-phase times of single flat k-means points (which part of the loop a sweep point spends its time in):
   python tools/micro/point_kmeans.py N D K [N D K ...]"""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from hsg_b200 import _lib
from hsg_b200.utils.segsort import common as S

dev = torch.device('cuda:0')
lib = _lib.load()
args = [int(a) for a in sys.argv[1:]] or [10000000, 512, 2048, 10000000, 256, 2048]
T = 10
for i in range(0, len(args), 3):
  nn, d, k = args[i:i + 3]
  g = torch.Generator(device=dev); g.manual_seed(235)
  x = torch.randn(nn, d, device=dev, generator=g); x = x / x.norm(dim=1, keepdim=True)
  init = torch.randint(0, k, (nn,), device=dev, generator=g)
  S.kmeans_with_initial_labels(x, init, k, 2); torch.cuda.synchronize()
  lib.hsg_profile_enable(1)
  S.kmeans_with_initial_labels(x, init, k, T); torch.cuda.synchronize()
  tot = np.zeros(len(bench.PHASES)); cnt = np.zeros(len(bench.PHASES), dtype=np.int64)
  lib.hsg_profile_collect(tot.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                          cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), len(bench.PHASES))
  lib.hsg_profile_enable(0)
  print(nn, d, k, ' '.join('%s=%.2f' % (bench.PHASES[j], tot[j] / T) for j in range(len(bench.PHASES)) if cnt[j]), '(ms per iteration)', flush=True)
  del x, init; torch.cuda.empty_cache()
