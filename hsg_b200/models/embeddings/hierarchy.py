"""Batched forms of the per-image(-pair) prototype bookkeeping of the HSG embedding
model (hsg/models/embeddings/resnet_fcn_hsg.py):

  calculate_kmeans_prototypes                  _calculate_kmeans_prototypes, :455-577 (one view)
                                               and :1005-1136 (image + augmented view)
  collect_nd_coarser_prototype                 _collect_nd_coarser_prototype, :683-748
  collect_pixel_hierarchical_clustering_indices  :751-780

The reference walks the images in a Python loop (`nonzero`, `gather`, `index_select`,
`unique`, `scatter_add_` per image: two host synchronisations and ~25 launches each).
Here every image is handled by the same launches: the dense prototype ids that
segment_by_kmeans produced are ranked per image group on P-sized tensors, the two
poolings are one deterministic segmented reduction each (K3, `ops.segment_reduce`),
and the per-pixel outputs are gathers.  Two small host reads remain (the number of
prototype ids, and the number of image groups / largest group, which size the outputs).

`patch_methods()` rebinds the three methods on the reference's model classes; they
keep their signatures and return tuples.
"""

import torch

from ... import _lib, ops
from ..._lib import HsgError


def _scatter_first(values, index, size, fill=0):
  """out[index[i]] = values[i] where every writer of a slot carries the same value."""
  out = torch.full((size,), fill, dtype=values.dtype, device=values.device)
  out.scatter_(0, index, values)
  return out


def calculate_kmeans_prototypes(cluster_embeddings, cluster_indices, cluster_batch_indices,
                                cluster_pos_embeddings, cluster_labels, image_indices=None,
                                label_divisor=2048, max_num_clusters=256):
  """Returns (prototypes [G,C,M], pos_prototypes [G,C,M] or None, padding_masks [G,M] bool,
  prototype_labels [G,M], prototype_batch_indices [G,M] (pad -1), cluster_indices_by_image [N]),
  G = image groups in increasing id order, M = max_num_clusters (reference :1005-1136), or the largest
  number of prototypes of a group when max_num_clusters is None (the `_cs` model).
  `cluster_indices` are the dense ids of segment_by_kmeans (one id = one (image, cluster,
  label) triple), as in the reference's call sites (:243-252, :906-916)."""
  emb = cluster_embeddings
  dev = emb.device
  cidx = cluster_indices.reshape(-1).long()
  bidx = cluster_batch_indices.reshape(-1).long()
  labs = cluster_labels.reshape(-1).long()
  n = cidx.shape[0]
  if n == 0:
    raise ValueError('calculate_kmeans_prototypes: no pixels')
  group_of_pixel = image_indices.to(dev).long()[bidx] if image_indices is not None else bidx
  grouped = (group_of_pixel[1:] >= group_of_pixel[:-1]).all() if n > 1 else torch.ones((), dtype=torch.bool, device=dev)
  p_all = int(cidx.max().item()) + 1                                    # host read 1: table size

  # prototype-level tables.  The reference ranks the distinct (cluster id, batch*div^2+label)
  # pairs of a group (prepare_prototype_labels, segsort/common.py:192-218); a cluster id of
  # segment_by_kmeans determines its batch index and label, so that is the rank of the
  # cluster id among the ids of its group.
  far = torch.iinfo(torch.int64).max
  p_group = _scatter_first(group_of_pixel, cidx, p_all, fill=far)       # absent ids sort last
  p_batch = _scatter_first(bidx, cidx, p_all, fill=-1)
  p_label = _scatter_first(labs, cidx, p_all, fill=-1)
  sorted_group, order = torch.sort(p_group, stable=True)                # prototypes by (group, id)
  valid_sorted = sorted_group != far
  is_new = torch.ones_like(valid_sorted)
  is_new[1:] = sorted_group[1:] != sorted_group[:-1]
  is_new &= valid_sorted
  pos_sorted = torch.arange(p_all, device=dev)
  group_rank_sorted = torch.cumsum(is_new.long(), 0) - 1                # 0..G-1
  start_of_group = torch.cummax(torch.where(is_new, pos_sorted, torch.zeros_like(pos_sorted)), 0).values
  local_sorted = pos_sorted - start_of_group                            # rank inside the group
  stats = torch.stack([is_new.sum(), torch.where(valid_sorted, local_sorted, torch.zeros_like(local_sorted)).max(),
                       grouped.long()]).tolist()                        # host read 2: output sizes
  g_count, biggest, grouped = int(stats[0]), int(stats[1]) + 1, bool(stats[2])
  # the Cityscapes model pads to the largest group of the batch (resnet_fcn_hsg_cs.py:499-502, 1061-1064)
  m = biggest if max_num_clusters is None else int(max_num_clusters)
  if biggest > m:
    raise HsgError('calculate_kmeans_prototypes: %d prototypes in one image group, max_num_clusters is %d '
                   '(the reference scatters out of bounds here, resnet_fcn_hsg.py:1090-1091)' % (biggest, m))
  slot_sorted = torch.where(valid_sorted, group_rank_sorted * m + local_sorted, torch.full_like(local_sorted, -1))
  slot_of_proto = torch.empty_like(slot_sorted)
  slot_of_proto[order] = slot_sorted
  local_of_proto = torch.empty_like(local_sorted)
  local_of_proto[order] = local_sorted

  slot_of_pixel = slot_of_proto[cidx]
  bins = g_count * m
  prototypes = ops.segment_reduce(emb, slot_of_pixel, bins, _lib.REDUCE_NORMALIZE)                  # :1090-1091
  prototypes = prototypes.view(g_count, m, -1).permute(0, 2, 1)
  pos_prototypes = None
  if cluster_pos_embeddings is not None:
    pos_prototypes = ops.segment_reduce(cluster_pos_embeddings, slot_of_pixel, bins, _lib.REDUCE_MEAN)  # :1115-1121
    pos_prototypes = pos_prototypes.view(g_count, m, -1).permute(0, 2, 1)

  flat_labels = torch.full((bins,), -1, dtype=torch.int64, device=dev)
  flat_batch = torch.full((bins,), -1, dtype=torch.int64, device=dev)
  ok = slot_of_proto >= 0
  flat_labels[slot_of_proto[ok]] = p_label[ok]
  flat_batch[slot_of_proto[ok]] = p_batch[ok]
  padding_masks = (flat_batch < 0).view(g_count, m)

  # pixels group by group, original order inside: the order the reference concatenates in
  by_image = local_of_proto[cidx]
  if not grouped:
    by_image = by_image[torch.sort(group_of_pixel, stable=True).indices]
  return (prototypes, pos_prototypes, padding_masks, flat_labels.view(g_count, m),
          flat_batch.view(g_count, m), by_image)


def collect_nd_coarser_prototype(prototypes, prototype_grouping_labels, prototype_padding_masks=None,
                                 num_groups=None, normalized=True):
  """Masked mean of the node representations of each coarser group ([B,C,N] -> [B,C,G]),
  reference :683-748; the atomics of its two scatter_add_ calls become one deterministic
  segmented reduction."""
  from ...utils.general import common as g_common
  b, c, nodes = prototypes.shape
  labels = prototype_grouping_labels.long()
  if num_groups is None:
    num_groups = int(labels.max().item()) + 1
  num_groups = int(num_groups)
  if prototype_padding_masks is not None:
    labels = labels.masked_fill(prototype_padding_masks, num_groups)       # dummy bin for padded nodes
  rows = prototypes.permute(0, 2, 1).reshape(b * nodes, c)
  bins = labels + torch.arange(b, device=labels.device).view(b, 1) * (num_groups + 1)
  out = ops.segment_reduce(rows, bins.reshape(-1), b * (num_groups + 1), _lib.REDUCE_MEAN)
  out = out.view(b, num_groups + 1, c)[:, :-1, :]
  if normalized:
    out = g_common.normalize_embedding(out)
  return out.permute(0, 2, 1)


def collect_pixel_hierarchical_clustering_indices(cluster_indices_by_batch, cluster_batch_indices,
                                                   finehrchy_prototype_grouping_labels):
  """Per pixel: the grouping label of its prototype (reference :751-780).  Row i of the label
  table belongs to the i-th distinct batch index; pixels come out batch index by batch index."""
  bidx = cluster_batch_indices.reshape(-1).long()
  cidx = cluster_indices_by_batch.reshape(-1).long()
  n = bidx.shape[0]
  sorted_b, perm = torch.sort(bidx, stable=True)
  is_new = torch.ones_like(sorted_b, dtype=torch.bool)
  if n > 1:
    is_new[1:] = sorted_b[1:] != sorted_b[:-1]
  row = torch.cumsum(is_new.long(), 0) - 1
  return finehrchy_prototype_grouping_labels[row, cidx[perm]]


# --------------------------------------------------------------------------- drop-in methods
def _calculate_kmeans_prototypes_single(self, cluster_embeddings, cluster_indices, cluster_batch_indices,
                                        cluster_pos_embeddings, cluster_labels):
  return calculate_kmeans_prototypes(cluster_embeddings, cluster_indices, cluster_batch_indices,
                                     cluster_pos_embeddings, cluster_labels, None,
                                     self.label_divisor, self.max_num_clusters)


def _calculate_kmeans_prototypes_multiview(self, cluster_embeddings, cluster_indices, cluster_batch_indices,
                                           cluster_pos_embeddings, cluster_labels, image_indices):
  return calculate_kmeans_prototypes(cluster_embeddings, cluster_indices, cluster_batch_indices,
                                     cluster_pos_embeddings, cluster_labels, image_indices,
                                     self.label_divisor, self.max_num_clusters)


def _collect_nd_coarser_prototype(self, prototypes, prototype_grouping_labels, prototype_padding_masks=None,
                                  num_groups=None, normalized=True):
  return collect_nd_coarser_prototype(prototypes, prototype_grouping_labels, prototype_padding_masks,
                                      num_groups, normalized)


def _collect_pixel_hierarchical_clustering_indices(self, cluster_indices_by_batch, cluster_batch_indices,
                                                   finehrchy_prototype_grouping_labels):
  return collect_pixel_hierarchical_clustering_indices(cluster_indices_by_batch, cluster_batch_indices,
                                                       finehrchy_prototype_grouping_labels)


def _calculate_kmeans_prototypes_single_cs(self, cluster_embeddings, cluster_indices, cluster_batch_indices,
                                           cluster_pos_embeddings, cluster_labels):
  """resnet_fcn_hsg_cs.py:455-560: pads to the largest per-image cluster count of the batch."""
  return calculate_kmeans_prototypes(cluster_embeddings, cluster_indices, cluster_batch_indices,
                                     cluster_pos_embeddings, cluster_labels, None, self.label_divisor, None)


def _calculate_kmeans_prototypes_multiview_cs(self, cluster_embeddings, cluster_indices, cluster_batch_indices,
                                              cluster_pos_embeddings, cluster_labels, image_indices):
  """resnet_fcn_hsg_cs.py:1010-1135: pads to the largest per-image-pair cluster count of the batch."""
  return calculate_kmeans_prototypes(cluster_embeddings, cluster_indices, cluster_batch_indices,
                                     cluster_pos_embeddings, cluster_labels, image_indices, self.label_divisor, None)


METHODS = {
    'ResnetFcn': {
        '_calculate_kmeans_prototypes': _calculate_kmeans_prototypes_single,
        '_collect_nd_coarser_prototype': _collect_nd_coarser_prototype,
        '_collect_pixel_hierarchical_clustering_indices': _collect_pixel_hierarchical_clustering_indices,
    },
    'MultiviewResnetFcn': {
        '_calculate_kmeans_prototypes': _calculate_kmeans_prototypes_multiview,
    },
}

# hsg/models/embeddings/resnet_fcn_hsg_cs.py (Cityscapes): same classes, prototypes padded to the batch's largest group
METHODS_CS = {
    'ResnetFcn': dict(METHODS['ResnetFcn'], _calculate_kmeans_prototypes=_calculate_kmeans_prototypes_single_cs),
    'MultiviewResnetFcn': {
        '_calculate_kmeans_prototypes': _calculate_kmeans_prototypes_multiview_cs,
    },
}
