"""Who waits for whom inside the single-pass tensor-core E-step (estep_tc1_kernel): per role, the share
of its lifetime spent in each of its waits, averaged over the CTAs.

    HSG_TC_EXP=2 python tools/estep_timeline.py [--images 48]

producer : wait0 = ring full (waiting for the MMA warp to release a stage)
MMA      : wait0 = ring empty (waiting for pixel slabs to land), wait1 = accumulator not drained yet
epilogue : wait0 = accumulator not ready, wait1 = the sweep itself (TMEM loads + max + hit bits)
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from hsg_b200 import ops, _lib  # noqa: E402
from hsg_b200.utils.segsort import common as S  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--images', type=int, default=48)
  ap.add_argument('--size', type=int, default=448)
  ap.add_argument('--dim', type=int, default=256)
  ap.add_argument('--grid', type=int, default=16)
  ap.add_argument('--dist', default='iid')
  a = ap.parse_args()
  assert int(os.environ.get('HSG_TC_EXP', '0')) & 2, 'run with HSG_TC_EXP=2 (other experiment bits may be added)'
  dev = torch.device('cuda:0')
  emb = bench.make_embeddings(torch, a, dev, 235)
  ex = S.segment_by_kmeans_ex(emb, None, [a.grid, a.grid], iterations=0)
  del emb
  x = ex['embeddings_with_loc']
  k = a.grid ** 2
  off = ex['seg_offsets']
  lab = S._grid_init([a.grid, a.grid], (a.size, a.size), dev)[0].repeat(a.images)
  xh, xerr = ops.make_half_copy(x, a.dim)
  lib = _lib.load()
  sms = lib.hsg_device_sms()
  clk = torch.zeros((sms, 4, 3), dtype=torch.int64, device=dev)
  for it in range(3):
    cent = ops.kmeans_mstep(x, lab, k, seg_offsets=off, max_seg_len=a.size * a.size)
    lib.hsg_debug_set_tc_clock(ctypes.c_void_p(clk.data_ptr()))
    lab = ops.kmeans_estep(x, cent, seg_offsets=off, max_seg_len=a.size * a.size, xh=xh, xerr=xerr,
                           flags=_lib.KMEANS_FORCE_TC)
    torch.cuda.synchronize()
    lib.hsg_debug_set_tc_clock(None)
    c = clk.double().cpu()
    tiles = x.shape[0] / 128 / sms
    print('iteration %d: %.0f cycles per CTA, %.0f per 128-pixel tile' % (it + 1, float(c[:, 1, 0].mean()), float(c[:, 1, 0].mean()) / tiles))
    for r, name in enumerate(['producer', 'MMA issuer', 'epilogue 0', 'epilogue 1']):
      tot = c[:, r, 0].mean()
      print('  %-11s wait0 %5.1f %%   wait1 %5.1f %%   (per tile: %.0f / %.0f cycles)' % (
          name, 100 * float(c[:, r, 1].mean() / tot), 100 * float(c[:, r, 2].mean() / tot),
          float(c[:, r, 1].mean()) / tiles * (2 if r >= 2 else 1), float(c[:, r, 2].mean()) / tiles * (2 if r >= 2 else 1)))


if __name__ == '__main__':
  main()
