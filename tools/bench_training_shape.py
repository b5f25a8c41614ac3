"""Latency of the k-means stage at the reference's TRAINING shapes (COCO stage 2: 12 images of 28x28, D=128,
grid 4x4, 15 iterations; Cityscapes-like 16 x 48x48, D=256), where everything is launch-bound:
    python tools/bench_training_shape.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsg_b200 import ops, _lib
from hsg_b200.utils.segsort import common as S

dev = torch.device('cuda:0')
for (b, d, hw, grid, iters) in ((12, 128, 28, 4, 15), (32, 128, 14, 4, 15), (16, 256, 48, 4, 15)):
  g = torch.Generator(device=dev); g.manual_seed(235 + hw)
  emb = torch.randn(b, d, hw, hw, device=dev, generator=g)
  ids = S.segment_by_kmeans(emb, None, [grid, grid], iterations=iters)[3]
  check = int((ids * (torch.arange(ids.numel(), device=dev) % 1000003 + 1)).sum())      # same with HSG_KMEANS_NO_PERSISTENT=1
  for _ in range(3):
    S.segment_by_kmeans(emb, None, [grid, grid], iterations=iters)
  torch.cuda.synchronize()
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  lib = _lib.load()
  n0 = lib.hsg_launch_count()
  import time
  t0.record()
  reps = 20
  h0 = time.perf_counter()
  for _ in range(reps):
    S.segment_by_kmeans(emb, None, [grid, grid], iterations=iters)
  h1 = time.perf_counter()
  t1.record(); torch.cuda.synchronize()
  print('segment_by_kmeans %2d x %dx%d, D=%d, K=%d, T=%d: %.3f ms per call (host side %.3f ms), %d kernel launches of the library, cluster-id checksum %d'
        % (b, hw, hw, d, grid * grid, iters, t0.elapsed_time(t1) / reps, (h1 - h0) * 1e3 / reps, (lib.hsg_launch_count() - n0) // reps, check))

# the k-means call alone (one C call: prep, relabel and the Python around them excluded)
print('hsg_kmeans_f32 alone:')
for (b, d, hw, grid, iters) in ((12, 128, 28, 4, 15), (32, 128, 14, 4, 15), (16, 256, 48, 4, 15)):
  g = torch.Generator(device=dev); g.manual_seed(235 + hw)
  emb = torch.randn(b, d, hw, hw, device=dev, generator=g)
  ex = S.segment_by_kmeans_ex(emb, None, [grid, grid], iterations=0)
  x, off = ex['embeddings_with_loc'], ex['seg_offsets']
  init = S._grid_init([grid, grid], (hw, hw), dev)[0].repeat(b)
  for _ in range(3):
    ops.kmeans(x, init, grid * grid, iters, seg_offsets=off, max_seg_len=hw * hw)
  torch.cuda.synchronize()
  import time
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  c0 = time.perf_counter()
  t0.record()
  for _ in range(50):
    ops.kmeans(x, init, grid * grid, iters, seg_offsets=off, max_seg_len=hw * hw)
  t1.record()
  c1 = time.perf_counter()
  torch.cuda.synchronize()
  print('  %2d x %dx%d, D=%d: %.3f ms per call on the GPU, %.3f ms of host time per call' % (b, hw, hw, d, t0.elapsed_time(t1) / 50, (c1 - c0) * 1e3 / 50))
