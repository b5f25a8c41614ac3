"""Drop-in for hsg/utils/segsort/eval.py: top-k retrieval accuracy.

The reference sorts every row of the [N,P] affinity matrix (`argsort`, :32-34) to read
k columns; it runs every training step on the prototypes themselves
(hsg/models/predictions/hsg.py:113-118).  Here one kernel keeps a running top-k per row
over register-tiled fp32 products (hsg_topk_affinity_f32): the [N,P] matrix never exists.
"""

import torch

from ... import ops


def top_k_ranking(embeddings, labels, prototypes, prototype_labels, top_k=3):
  """(accuracy, retrieved labels [N,top_k]); reference :9-52."""
  embeddings = embeddings.reshape(-1, embeddings.shape[-1])
  prototypes = prototypes.reshape(-1, prototypes.shape[-1])
  if top_k <= 8 and top_k <= prototypes.shape[0]:
    top = ops.topk_affinity(embeddings, prototypes, top_k)           # one kernel, no [N,P] matrix (CUDA only)
  else:                                                              # k > 8 (retrieval, top-20): library GEMM + top-k
    affinity = torch.mm(embeddings, prototypes.t())
    top = torch.topk(affinity, top_k, dim=1, largest=True, sorted=True).indices
  retrieved = prototype_labels.reshape(-1)[top.reshape(-1)].view(-1, top_k)
  accuracy = torch.mean((retrieved == labels.reshape(-1, 1)).float())
  return accuracy, retrieved


def majority_label_from_topk(top_k_labels, num_classes=None):
  """Most frequent label of each row of [num_queries, top_k], ties -> lowest label (reference :55-72,
  which materialises the [N, k, C] one-hot tensor and sums it; here the counts are one scatter_add)."""
  labels = top_k_labels.reshape(top_k_labels.shape[0], -1).long()
  if num_classes is None:
    num_classes = int(labels.max()) + 1
  counts = torch.zeros((labels.shape[0], int(num_classes)), dtype=torch.long, device=labels.device)
  counts.scatter_add_(1, labels, torch.ones_like(labels))
  return torch.argmax(counts, 1)
