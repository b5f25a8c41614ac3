"""Drop-in operators for hsg/utils/general/common.py (the two on the hot path).

Same names, argument meaning, return dtypes and autograd behaviour as the
reference; the work happens in libhsgb200.so.
"""

import torch

from ... import ops
from ..._lib import REDUCE_MEAN


def normalize_embedding(embeddings, eps=1e-12):
  """L2-normalise the last dimension (reference :101-120).

  ``eps`` is fixed at the reference's default 1e-12 (the only value any caller
  uses); another value raises."""
  if eps != 1e-12:
    raise ValueError('normalize_embedding: only eps=1e-12 is supported')
  return ops.normalize(embeddings)


def segment_mean(x, index):
  """tf.segment_mean (reference :123-147): float32 [index.max()+1, C]."""
  index = index.reshape(-1)
  num = int(index.max()) + 1            # same host sync as the reference (:128)
  return ops.segment_reduce(x.reshape(-1, x.shape[-1]), index, num, REDUCE_MEAN)
