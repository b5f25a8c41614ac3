"""Golden vectors of the Cityscapes model's `_calculate_kmeans_prototypes`
(hsg/models/embeddings/resnet_fcn_hsg_cs.py:455-560 and :1010-1135: prototypes padded to the largest
per-image(-pair) cluster count of the batch instead of 256), on the inputs of tests/golden/kmeans_prototypes.npz.

    python oracle/gen_golden_cs.py        # build container only (/root/reference)

Test infrastructure: runs the UNMODIFIED reference on CPU and writes tests/golden/kmeans_prototypes_cs.npz.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('HSG_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, '..', 'tests', 'refstubs'))
GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')


def main():
  from hsg.models.embeddings.resnet_fcn_hsg_cs import MultiviewResnetFcn, ResnetFcn
  g = dict(np.load(os.path.join(GOLDEN, 'kmeans_prototypes.npz')))
  t = torch.from_numpy
  fake_self = types.SimpleNamespace(label_divisor=2048, max_num_clusters=256)
  mv = MultiviewResnetFcn._calculate_kmeans_prototypes(fake_self, t(g['emb']), t(g['cluster']), t(g['batch']), t(g['pos']),
                                                       t(g['labels']), t(g['image_indices']))
  sv = ResnetFcn._calculate_kmeans_prototypes(fake_self, t(g['emb']), t(g['cluster']), t(g['batch']), t(g['pos']), t(g['labels']))
  out = {}
  for prefix, res in (('mv', mv), ('sv', sv)):
    for i, v in enumerate(res):
      out['%s%d' % (prefix, i)] = v.detach().cpu().numpy()
  path = os.path.join(GOLDEN, 'kmeans_prototypes_cs.npz')
  np.savez_compressed(path, **out)
  print(path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
  main()
