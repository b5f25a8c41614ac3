"""Oracle: prototype pooling per image(-pair) and the cross-GPU gather step
(numpy restatement, test infrastructure).

Restates ``MultiviewResnetFcn._calculate_kmeans_prototypes``
(hsg/models/embeddings/resnet_fcn_hsg.py:1005-1136; single view :455-577) and
hsg/models/utils.py:41-240 of the reference.  "Lists of per-GPU tensors"
become lists of per-rank numpy arrays.
"""

import numpy as np

from . import ops


# --------------------------------------------------------------------------
# a9  _calculate_kmeans_prototypes  resnet_fcn_hsg.py:1005-1136 / :455-577
# --------------------------------------------------------------------------
def calculate_kmeans_prototypes(cluster_embeddings, cluster_indices,
                                cluster_batch_indices, cluster_pos_embeddings,
                                cluster_labels, image_indices=None,
                                label_divisor=2048, max_num_clusters=256):
  """Per image (or per image pair when ``image_indices`` maps batch index ->
  image id): dense re-index of the clusters, normalised prototype sums padded
  to ``max_num_clusters``, segment-mean of the positional embeddings, padding
  mask, prototype labels / batch indices (pad -1).

  Returns (prototypes [G,C,M], pos_prototypes [G,C,M] or None, padding_mask
  [G,M] bool, prototype_labels [G,M], prototype_batch_indices [G,M],
  cluster_indices_by_image [N])."""
  emb = np.asarray(cluster_embeddings, np.float32)
  cidx = np.asarray(cluster_indices, np.int64)
  bidx = np.asarray(cluster_batch_indices, np.int64)
  labs = np.asarray(cluster_labels, np.int64)
  if image_indices is not None:                                       # :1057-1060
    img_of_pixel = np.asarray(image_indices, np.int64)[bidx]
  else:
    img_of_pixel = bidx
  m = max_num_clusters
  div2 = int(label_divisor) ** 2
  protos, pos_protos, plabs, pbats, masks, by_image = [], [], [], [], [], []
  for img in np.unique(img_of_pixel):                                 # :1074
    sel = np.nonzero(img_of_pixel == img)[0]
    c_labs = bidx[sel] * div2 + labs[sel]                             # :1079-1080
    proto_labs, c_inds = ops.prepare_prototype_labels(
        c_labs, cidx[sel], int(c_labs.max()) + 1)                     # :1082-1083
    proto_bat = proto_labs // div2
    proto_labs = proto_labs % div2
    n = proto_labs.shape[0]
    if n > m:
      raise IndexError('more than max_num_clusters prototypes in one image group '
                       '(the reference scatters out of bounds here, :1090-1091)')
    protos.append(ops.calculate_prototypes_from_labels(emb[sel], c_inds, m))   # :1090-1091
    plabs.append(np.pad(proto_labs, (0, m - n), constant_values=-1))
    pbats.append(np.pad(proto_bat, (0, m - n), constant_values=-1))
    masks.append(np.arange(m) >= n)                                   # :1100-1105
    by_image.append(c_inds)
    if cluster_pos_embeddings is not None:                            # :1115-1121
      pm = ops.segment_mean(np.asarray(cluster_pos_embeddings, np.float32)[sel], c_inds)
      pos_protos.append(np.pad(pm, ((0, m - n), (0, 0))))
  protos = np.transpose(np.stack(protos, 0), (0, 2, 1))               # [G,C,M]
  pos = np.transpose(np.stack(pos_protos, 0), (0, 2, 1)) if pos_protos else None
  return (protos, pos, np.stack(masks, 0), np.stack(plabs, 0),
          np.stack(pbats, 0), np.concatenate(by_image, 0))


# --------------------------------------------------------------------------
# a13  cross-GPU gathers            hsg/models/utils.py:41-240
# --------------------------------------------------------------------------
def gather_clustering_and_update_prototypes(embeddings, embeddings_with_loc,
                                            cluster_indices, batch_indices,
                                            semantic_labels, instance_labels):
  """utils.py:127-217.  Inputs are lists (one entry per rank).  Returns global
  (prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch) plus a
  list with each rank's updated pixel -> prototype ids."""
  sizes = [np.shape(c)[0] for c in cluster_indices]
  emb = np.concatenate([np.asarray(e, np.float32) for e in embeddings], 0)
  emb_loc = np.concatenate([np.asarray(e, np.float32) for e in embeddings_with_loc], 0)
  cidx = np.concatenate([np.asarray(c, np.int64) for c in cluster_indices], 0)
  bidx = np.concatenate([np.asarray(c, np.int64) for c in batch_indices], 0)
  sem = np.concatenate([np.asarray(c, np.int64) for c in semantic_labels], 0)
  inst = np.concatenate([np.asarray(c, np.int64) for c in instance_labels], 0)

  divisor = int(cidx.max()) + 1                                       # :181
  _, cidx = np.unique(bidx * divisor + cidx, return_inverse=True)     # :182-183
  cidx = cidx.reshape(-1)
  lab_div = max(int(inst.max()) + 1, int(sem.max()) + 1)              # :186
  labels = bidx * lab_div ** 2 + sem * lab_div + inst                 # :187-189
  proto_labels, upd = ops.prepare_prototype_labels(labels, cidx, int(labels.max()) + 1)
  proto_batch = proto_labels // lab_div ** 2                          # :195-197
  proto_sem = (proto_labels % lab_div ** 2) // lab_div
  proto_inst = proto_labels % lab_div
  prototypes = ops.calculate_prototypes_from_labels(emb, upd)         # :199-202
  prototypes_with_loc = ops.calculate_prototypes_from_labels(emb_loc, upd)
  split = np.split(upd, np.cumsum(sizes)[:-1])
  return prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch, split


def gather_and_update_cluster_mappings(cluster_indices_1, cluster_indices_2):
  """utils.py:78-124: table mapping ids of level 1 -> ids of level 2."""
  c1 = np.concatenate([np.asarray(c, np.int64) for c in cluster_indices_1], 0)
  c2 = np.concatenate([np.asarray(c, np.int64) for c in cluster_indices_2], 0)
  max_ind = int(c2.max()) + 1
  mapping = np.unique(c1 * max_ind + c2)
  table = np.zeros(int(c1.max()) + 1, np.int64)
  table[mapping // max_ind] = mapping % max_ind
  return table


def gather_and_reorder_image_indices(image_indices):
  """utils.py:41-74: relabel image ids by first occurrence over all ranks.
  Returns the FULL concatenated vector (the reference hands every GPU the whole
  vector, :72; it is indexed with batch indices that carry the +B*gpu offset)."""
  ids = np.concatenate([np.asarray(c, np.int64) for c in image_indices], 0)
  _, inv = np.unique(ids, return_inverse=True)
  inv = inv.reshape(-1)
  first = np.full(int(inv.max()) + 1, len(inv), np.int64)
  np.minimum.at(first, inv, np.arange(len(inv)))
  _, out = np.unique(first[inv], return_inverse=True)
  return out.reshape(-1).astype(np.int64)


# --------------------------------------------------------------------------
# a12  hierarchy helpers            resnet_fcn_hsg.py:683-780
# --------------------------------------------------------------------------
def collect_nd_coarser_prototype(prototypes, grouping_labels, padding_masks=None,
                                 num_groups=None, normalized=True):
  """resnet_fcn_hsg.py:683-748: per batch entry, mean of the node columns of each
  coarser group (padded nodes go to a dummy group that is dropped), optional L2
  normalisation.  [B,C,N] -> [B,C,G]."""
  p = np.asarray(prototypes, np.float32)
  lab = np.asarray(grouping_labels, np.int64).copy()
  b, c, nodes = p.shape
  if num_groups is None:
    num_groups = int(lab.max()) + 1                                   # :707-708
  if padding_masks is not None:
    lab[np.asarray(padding_masks, bool)] = num_groups                 # :716-719
  out = np.zeros((b, num_groups + 1, c), np.float32)
  cnt = np.zeros((b, num_groups + 1, c), np.float32)
  for i in range(b):
    np.add.at(out[i], lab[i], p[i].T)                                 # :733-737
    np.add.at(cnt[i], lab[i], np.ones((nodes, c), np.float32))
  out = out / np.maximum(cnt, np.float32(1e-12))                      # :738-739
  out = out[:, :-1, :]
  if normalized:
    out = ops.normalize_embedding(out)                                # :743-744
  return np.transpose(out, (0, 2, 1))


def collect_pixel_hierarchical_clustering_indices(cluster_indices_by_batch,
                                                  cluster_batch_indices, grouping_labels):
  """resnet_fcn_hsg.py:751-780: row i of `grouping_labels` belongs to the i-th
  distinct batch index; pixels are emitted batch index by batch index."""
  cidx = np.asarray(cluster_indices_by_batch, np.int64)
  bidx = np.asarray(cluster_batch_indices, np.int64)
  lab = np.asarray(grouping_labels, np.int64)
  out = []
  for i, b in enumerate(np.unique(bidx)):
    out.append(lab[i][cidx[bidx == b]])
  return np.concatenate(out, 0)
