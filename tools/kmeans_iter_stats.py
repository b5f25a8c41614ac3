"""Per-iteration anatomy of the k-means loop at the benchmark shape: how many pixels the tensor-core
pass sends to the float64 re-decision (and how many of those scan every cluster), and the
CUDA-event time of every phase of that iteration (library phase profiler).

    python tools/kmeans_iter_stats.py [--images 48] [--dist iid|planted] [--iters 10]

Drives the loop one step at a time through the single-step entry points (hsg_kmeans_mstep_f32 /
hsg_kmeans_estep_f32), so the M-step here is always a full pass; the E-step is the production one.
"""
import argparse
import ctypes
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402  (make_embeddings, PHASES)
from hsg_b200 import ops, _lib  # noqa: E402
from hsg_b200.utils.segsort import common as S  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--images', type=int, default=48)
  ap.add_argument('--size', type=int, default=448)
  ap.add_argument('--dim', type=int, default=256)
  ap.add_argument('--grid', type=int, default=16)
  ap.add_argument('--iters', type=int, default=10)
  ap.add_argument('--dist', default='iid')
  a = ap.parse_args()
  dev = torch.device('cuda:0')
  emb = bench.make_embeddings(torch, a, dev, 235)
  ex = S.segment_by_kmeans_ex(emb, None, [a.grid, a.grid], iterations=0)
  del emb
  x = ex['embeddings_with_loc']
  n = x.shape[0]
  k = a.grid ** 2
  off = ex['seg_offsets']
  lab = S._grid_init([a.grid, a.grid], (a.size, a.size), dev)[0].repeat(a.images)
  xh, xerr = ops.make_half_copy(x, a.dim)
  lib = _lib.load()
  names = bench.PHASES
  print('# %d pixels, D=%d, K=%d, %s' % (n, a.dim, k, a.dist))
  print('# iter  changed   listed   (frac)   scan_all   ' + '  '.join('%s_ms' % p for p in ('mstep', 'convert', 'estep', 'estep_fixup')))
  for it in range(a.iters):
    lib.hsg_profile_enable(1)
    cent = ops.kmeans_mstep(x, lab, k, seg_offsets=off, max_seg_len=a.size * a.size)
    new, nre = ops.kmeans_estep(x, cent, seg_offsets=off, max_seg_len=a.size * a.size, xh=xh, xerr=xerr,
                                flags=_lib.KMEANS_FORCE_TC, return_rechecked=True)
    torch.cuda.synchronize()
    tot = (ctypes.c_double * len(names))()
    cnt = (ctypes.c_longlong * len(names))()
    lib.hsg_profile_collect(tot, cnt, len(names))
    lib.hsg_profile_enable(0)
    ms = dict(zip(names, list(tot)))
    changed = float((new != lab).float().mean())
    listed, scan_all = int(nre[0]), int(nre[1])
    print('%5d  %.4f  %9d  %.4f  %9d   %.3f  %.3f  %.3f  %.3f' % (
        it + 1, changed, listed, listed / n, scan_all,
        ms['mstep_sort'] + ms['mstep_gather'] + ms['mstep_combine'], ms['convert'], ms['estep'], ms['estep_fixup']))
    lab = new


if __name__ == '__main__':
  main()
