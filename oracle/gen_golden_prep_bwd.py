"""Golden vectors for the backward of the prep chain: the reference's autograd through
segment_by_kmeans' permute / normalize / cat(loc) / normalize / index_select
(hsg/utils/segsort/common.py:305-365) at D = 256, with dropped (ignored) pixels and a zero pixel.

    python oracle/gen_golden_prep_bwd.py      # build container only (/root/reference)

Test infrastructure: runs the UNMODIFIED reference on CPU (one external CPU shim, see tests/refenv.py)
and writes tests/golden/prep_backward.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', 'tests'))
import refenv  # noqa: E402

refenv.REF = os.environ.get('HSG_REFERENCE', '/root/reference')


def main():
  sys.path.insert(0, refenv.STUBS)
  sys.path.insert(0, refenv.REF)
  seg = refenv.cpu_segment_by_kmeans()
  torch.manual_seed(235)
  b, d, h, w = 2, 256, 6, 10
  emb = torch.randn(b, d, h, w)
  emb[1, :, 2, 3] = 0.0                                   # a zero row: both normalisations take the eps branch
  emb.requires_grad_(True)
  labels = torch.randint(0, 3, (b, h, w))
  labels[0, 0, :4] = 7
  labels[1, 5, 9] = 7
  x, xloc, lab, cidx, bidx = seg(emb, labels, [2, 2], ignore_index=7, iterations=1)
  gx = torch.randn_like(x)
  gz = torch.randn_like(xloc)
  ((x * gx).sum() + (xloc * gz).sum()).backward()
  both = emb.grad.clone()
  emb.grad = None
  x, xloc, _, _, _ = seg(emb, labels, [2, 2], ignore_index=7, iterations=1)
  (x * gx).sum().backward()
  only_x = emb.grad.clone()
  out = os.path.join(HERE, '..', 'tests', 'golden', 'prep_backward.npz')
  np.savez_compressed(out, emb=emb.detach().numpy(), labels=labels.numpy(), gx=gx.numpy(), gz=gz.numpy(),
                      x=x.detach().numpy(), xloc=xloc.detach().numpy(), demb=both.numpy(), demb_x_only=only_x.numpy())
  print(out, os.path.getsize(out))


if __name__ == '__main__':
  main()
