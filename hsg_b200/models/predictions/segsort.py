"""Nearest-neighbour label retrieval of the SegSort prediction head
(hsg/models/predictions/segsort.py, `Segsort.predictions` :66-123), as used by
pyscripts/inference/inference.py:220-224.

The reference pools one prototype per k-means cluster of the image, then walks the prototypes in
ten groups, each with a [n, M] affinity matrix, a full `argsort` of every row to read 20 columns
and a [n, 20, C] one-hot tensor for the vote.  Here: one pooling launch (K3), one [P, M] GEMM, a
top-20 selection and a scatter_add vote.  Same return tuple and dtypes.  (Two bank prototypes with
bit-identical affinity to a query may come out in a different order than `argsort` puts them.)
"""

import torch

from ..._lib import HsgError
from ...utils.segsort import common as segsort_common
from ...utils.segsort import eval as segsort_eval

TOP_K = 20      # reference :99, :109


def predictions(self, datas, targets={}):
  """(semantic_pred [N] int64, semantic_topk [N, 20] int64) per pixel, or (None, None) when the
  memory bank or the clustering outputs are missing (reference :70-84)."""
  memory_prototypes = targets.get('semantic_memory_prototype', None)
  memory_labels = targets.get('semantic_memory_prototype_label', None)
  cluster_embeddings = datas.get('cluster_embedding', None)
  cluster_indices = datas.get('cluster_index', None)
  if memory_prototypes is None or memory_labels is None or cluster_embeddings is None or cluster_indices is None:
    return None, None
  if memory_prototypes.shape[0] < TOP_K:
    raise HsgError('predictions: the memory bank holds %d prototypes, the vote reads the top %d'
                   % (memory_prototypes.shape[0], TOP_K))
  dev = cluster_embeddings.device
  _, cluster_indices = torch.unique(cluster_indices.reshape(-1), return_inverse=True)        # :86-87
  num_prototypes = int(cluster_indices.max()) + 1
  prototypes = segsort_common.calculate_prototypes_from_labels(cluster_embeddings, cluster_indices, num_prototypes)
  _, top_k_labels = segsort_eval.top_k_ranking(
      prototypes, torch.zeros(num_prototypes, dtype=torch.long, device=dev),
      memory_prototypes.to(dev), memory_labels.to(dev), TOP_K)                               # :105-111
  majority = segsort_eval.majority_label_from_topk(top_k_labels)                             # :112-113
  return majority[cluster_indices], top_k_labels[cluster_indices]                            # :116-121


METHODS = {'Segsort': {'predictions': predictions}}
