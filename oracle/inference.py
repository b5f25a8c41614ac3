"""Oracle: prototype bank and nearest-neighbour label retrieval of the inference
scripts (numpy restatement, test infrastructure -- SURVEY section 8f rank 4).

Restates, with the reference file:line each function follows,
  generate_clusters           hsg/models/embeddings/resnet_fcn.py:90-148
  find_majority_label_index   hsg/utils/segsort/common.py:221-267
  prototype bank of one image pyscripts/inference/prototype.py:181-208
  load_memory_banks           hsg/utils/segsort/others.py:11-41
  majority_label_from_topk    hsg/utils/segsort/eval.py:55-72
  Segsort.predictions         hsg/models/predictions/segsort.py:66-123
Only tests/ (and bench.py's CPU legs) may import this package.
"""

import glob
import os

import numpy as np

from . import ops


def generate_clusters(embeddings, semantic_labels, instance_labels, label_divisor, semantic_ignore_index,
                      kmeans_num_clusters, kmeans_iterations, local_features=None):
  """resnet_fcn.py:90-148.  Returns the reference's dict with numpy values."""
  if semantic_labels is not None and instance_labels is not None:
    sem = np.asarray(semantic_labels, np.int64)
    labels = sem * label_divisor + np.asarray(instance_labels, np.int64)             # :113
    ignore_index = int(labels.max()) + 1                                             # :114
    labels = np.where(sem == semantic_ignore_index, ignore_index, labels)            # :115-117
  else:
    labels, ignore_index = None, None
  emb, emb_loc, lab, clu, bat = ops.segment_by_kmeans(embeddings, labels, kmeans_num_clusters,
                                                      local_features=local_features, ignore_index=ignore_index,
                                                      iterations=kmeans_iterations)   # :123-134
  return {'cluster_embedding': emb, 'cluster_embedding_with_loc': emb_loc,
          'cluster_semantic_label': lab // label_divisor, 'cluster_instance_label': lab % label_divisor,   # :136-137
          'cluster_index': clu, 'cluster_batch_index': bat}


def find_majority_label_index(semantic_labels, cluster_labels):
  """segsort/common.py:221-267: votes [num_clusters, num_classes]; majority = first arg-max; the pixels that
  carry their cluster's majority label, as an [n,1] index array (torch `nonzero` of a 1-D tensor)."""
  sem = np.asarray(semantic_labels, np.int64).reshape(-1)
  clu = np.asarray(cluster_labels, np.int64).reshape(-1)
  num_clusters, num_classes = int(clu.max()) + 1, int(sem.max()) + 1                 # :239-240
  votes = np.zeros((num_clusters, num_classes), np.int64)
  np.add.at(votes, (clu, sem), 1)                                                    # :250-258
  majority = votes.argmax(1)                                                         # :259
  keep = np.nonzero(majority[clu] == sem)[0].reshape(-1, 1)                          # :261-267
  return keep, majority


def prototype_bank(cluster_embeddings, cluster_indices, semantic_labels):
  """prototype.py:193-203: (prototypes [P,D], prototype_labels [P]) of one image."""
  prototypes = ops.calculate_prototypes_from_labels(cluster_embeddings, cluster_indices)
  _, prototype_labels = find_majority_label_index(semantic_labels, cluster_indices)
  return prototypes, prototype_labels


def load_memory_banks(memory_dir):
  """others.py:11-41: the *.npy dicts of a directory, in sorted file order, concatenated."""
  paths = sorted(glob.glob(os.path.join(memory_dir, '*.npy')))
  assert paths, 'No memory stored in the directory'
  datas = [np.load(p, allow_pickle=True).item() for p in paths]
  return (np.concatenate([d['prototype'] for d in datas], 0).astype(np.float32),
          np.concatenate([d['prototype_label'] for d in datas], 0).astype(np.int64))


def majority_label_from_topk(top_k_labels, num_classes=None):
  """eval.py:55-72: most frequent label of each row, ties -> lowest label."""
  lab = np.asarray(top_k_labels, np.int64)
  if num_classes is None:
    num_classes = int(lab.max()) + 1
  counts = np.zeros((lab.shape[0], num_classes), np.int64)
  np.add.at(counts, (np.repeat(np.arange(lab.shape[0]), lab.shape[1]), lab.reshape(-1)), 1)
  return counts.argmax(1)          # the reference sums the k one-hot rows of a query ([N,k,C] -> [N,C]) = these counts


def predictions(cluster_embeddings, cluster_indices, memory_prototypes, memory_prototype_labels, top_k=20):
  """segsort.py:66-123: prototypes of the image's clusters, their top-k most similar bank prototypes
  (descending affinity), majority label per cluster, scattered back to the pixels.  The reference walks
  the clusters in ten groups (:101-117); the grouping does not change any value."""
  _, clu = np.unique(np.asarray(cluster_indices).reshape(-1), return_inverse=True)   # :86-87
  clu = clu.reshape(-1)
  protos = ops.calculate_prototypes_from_labels(cluster_embeddings, clu, int(clu.max()) + 1)   # :89-93
  aff = protos.astype(np.float32) @ np.asarray(memory_prototypes, np.float32).T      # eval.py:31
  top = np.argsort(-aff, axis=1, kind='stable')[:, :top_k]                           # eval.py:32-33
  top_labels = np.asarray(memory_prototype_labels, np.int64)[top]                    # eval.py:45-49
  majority = majority_label_from_topk(top_labels)                                    # :112-113
  return majority[clu], top_labels[clu]                                              # :116-121
