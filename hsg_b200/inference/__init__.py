"""The steps either side of the hot path in the reference's inference scripts
(pyscripts/inference/prototype.py:181-208, inference.py:207-224):
per-image k-means on a full-resolution embedding, the prototype bank entry of the image, and the
nearest-neighbour labels against a loaded bank.  All three run on the same kernels as training
(K0-K3); nothing here falls back to the CPU.

    out = generate_clusters(embedding, fake_sem, fake_inst, label_divisor=2048, semantic_ignore_index=255,
                            kmeans_num_clusters=[12, 24], kmeans_iterations=10)
    prototypes, labels = prototype_bank(out['cluster_embedding'], out['cluster_index'], semantic_gt)
    save_memory_bank(path, prototypes, labels)
    ...
    bank = load_memory_banks(directory)
    pred, topk = nearest_neighbor_labels(out, *bank)
"""

from ..models.predictions.segsort import predictions as _predictions
from ..utils.segsort import common as segsort_common
from ..utils.segsort.others import load_memory_banks, save_memory_bank

__all__ = ['generate_clusters', 'prototype_bank', 'nearest_neighbor_labels', 'load_memory_banks', 'save_memory_bank']


def generate_clusters(embeddings, semantic_labels, instance_labels, label_divisor, semantic_ignore_index,
                      kmeans_num_clusters, kmeans_iterations, local_features=None):
  """`ResnetFcn.generate_clusters` (hsg/models/embeddings/resnet_fcn.py:90-148) as a function of its four
  config fields.  embeddings [B,C,H,W]; labels [B,H,W] int64 or both None.  Returns the reference's dict."""
  if semantic_labels is not None and instance_labels is not None:
    labels = semantic_labels * label_divisor + instance_labels
    ignore_index = labels.max() + 1
    labels = labels.masked_fill(semantic_labels == semantic_ignore_index, ignore_index)
  else:
    labels, ignore_index = None, None
  emb, emb_loc, lab, cidx, bidx = segsort_common.segment_by_kmeans(
      embeddings, labels, kmeans_num_clusters, local_features=local_features, ignore_index=ignore_index,
      iterations=kmeans_iterations)
  return {'cluster_embedding': emb, 'cluster_embedding_with_loc': emb_loc,
          'cluster_semantic_label': lab // label_divisor, 'cluster_instance_label': lab % label_divisor,
          'cluster_index': cidx, 'cluster_batch_index': bidx}


def prototype_bank(cluster_embeddings, cluster_indices, semantic_labels):
  """One image's bank entry (pyscripts/inference/prototype.py:193-203): the mean direction of every cluster
  and the majority ground-truth label of its pixels.  `semantic_labels` has one entry per clustered pixel."""
  prototypes = segsort_common.calculate_prototypes_from_labels(cluster_embeddings, cluster_indices)
  _, prototype_labels = segsort_common.find_majority_label_index(semantic_labels, cluster_indices)
  return prototypes, prototype_labels


def nearest_neighbor_labels(clustering_outputs, memory_prototypes, memory_prototype_labels):
  """(label [N], top-20 retrieved labels [N,20]) per pixel: `Segsort.predictions`
  (hsg/models/predictions/segsort.py:66-123) on the outputs of `generate_clusters`."""
  return _predictions(None, clustering_outputs, {'semantic_memory_prototype': memory_prototypes,
                                                 'semantic_memory_prototype_label': memory_prototype_labels})
