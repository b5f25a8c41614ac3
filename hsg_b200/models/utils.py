"""Drop-in operators for hsg/models/utils.py -- the cross-GPU step.

The reference (single process, one thread per GPU) concatenates EVERY pixel
embedding of every GPU on one anchor GPU, re-ranks the clusters there, pools
the prototypes and copies them back (hsg/models/utils.py:127-217): N*(D+D')*4
bytes through one device, three times per step.  Here each GPU pools the
prototypes of its own images (complete locally, because images never straddle
GPUs and batch indices are rank-major) and only the [P_r, D] prototypes and
their labels travel.  SURVEY.md A.3b: the results are identical.

Two call styles:
  * the reference's own list-of-per-GPU-tensors signatures (one process, any
    number of devices) -- what pyscripts/train/train.py calls;
  * ``dist_*`` variants for one process per GPU (torchrun): NCCL all-gather of
    the prototypes (backward = all-reduce of their gradients) and an all-reduce
    of the centroid sums inside flat k-means.
"""

import torch
import torch.distributed as dist

from .. import ops
from .._lib import REDUCE_NORMALIZE, REDUCE_SUM

_PACK = 1 << 31   # order-preserving packing of (semantic, instance) pairs


# ---------------------------------------------------------------- local stage (per GPU)
_RELABEL_TABLE_LIMIT = 1 << 28   # relabel.cu: presence table of num_images * kmax * n_label_values entries


def _rank_triples_by_sort(batch_indices, cluster_indices, packed):
  """Sort-based ranking of the distinct (batch, cluster, label) triples for the shapes whose presence table
  would not fit (many images x many dense cluster ids x many instance labels): two `torch.unique` passes on the
  device, which is what the reference itself does (utils.py:181-194, segsort/common.py:82-86).  Same ids and
  the same lexicographic order as the table path.  Returns (ids, proto_label, proto_batch, n)."""
  cdiv = cluster_indices.max() + 1
  groups, g = torch.unique(batch_indices * cdiv + cluster_indices, return_inverse=True)
  label_values, l = torch.unique(packed, return_inverse=True)
  nl = label_values.numel()
  keys, ids = torch.unique(g * nl + l, return_inverse=True)
  proto_label = label_values[keys % nl]
  proto_batch = torch.div(groups[torch.div(keys, nl, rounding_mode='floor')], cdiv, rounding_mode='floor')
  return ids, proto_label, proto_batch, keys.numel()


def _local_rank_and_pool(embeddings, embeddings_with_loc, cluster_indices, batch_indices,
                         semantic_labels, instance_labels):
  """Dense ids of the distinct (batch, cluster, sem, inst) tuples on this GPU
  (lexicographic, as utils.py:181-194 ranks them), the pooled prototypes and
  their labels."""
  dev = embeddings.device
  packed = semantic_labels.long() * _PACK + instance_labels.long()
  # host-side scalars, like the reference's .max() calls
  b0, b1 = int(batch_indices.min()), int(batch_indices.max())
  kmax = int(cluster_indices.max()) + 1
  label_values = torch.unique(packed)
  if (b1 - b0 + 1) * kmax * label_values.numel() >= _RELABEL_TABLE_LIMIT:
    ids, pl, pb, n = _rank_triples_by_sort(batch_indices.long(), cluster_indices.long(), packed)
  else:
    ids, pl, pb, _, npro = ops.relabel(batch_indices.long().contiguous(), cluster_indices.long().contiguous(),
                                       packed.contiguous(), b0, b1 - b0 + 1, kmax, label_values)
    n = int(npro)
  protos = ops.segment_reduce(embeddings, ids, n, REDUCE_NORMALIZE)
  protos_loc = ops.segment_reduce(embeddings_with_loc, ids, n, REDUCE_NORMALIZE)
  pl = pl[:n]
  return ids, protos, protos_loc, pl // _PACK, pl % _PACK, pb[:n], (b0, b1)


def _check_rank_major(ranges):
  for (a0, a1), (c0, c1) in zip(ranges, ranges[1:]):
    if c0 <= a1:
      raise RuntimeError('batch indices of different GPUs overlap or are not rank-major '
                         '(%d..%d then %d..%d): images must not straddle GPUs' % (a0, a1, c0, c1))


# ---------------------------------------------------------------- reference (list) signatures
def gather_clustering_and_update_prototypes(embeddings, embeddings_with_loc, cluster_indices,
                                            batch_indices, semantic_labels, instance_labels,
                                            anchor_device=None):
  """Reference utils.py:127-217, same lists in / lists out."""
  devices = [c.device for c in cluster_indices]
  local = [_local_rank_and_pool(*args) for args in zip(embeddings, embeddings_with_loc, cluster_indices,
                                                       batch_indices, semantic_labels, instance_labels)]
  _check_rank_major([l[6] for l in local])
  offsets, total = [], 0
  for l in local:
    offsets.append(total)
    total += l[1].shape[0]

  def everywhere(k):
    return [torch.cat([l[k].to(d) for l in local], 0) for d in devices]

  prototypes, prototypes_with_loc = everywhere(1), everywhere(2)
  proto_sem, proto_inst, proto_batch = everywhere(3), everywhere(4), everywhere(5)
  updated = [l[0] + off for l, off in zip(local, offsets)]
  return prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch, updated


def gather_and_update_cluster_mappings(cluster_indices_1, cluster_indices_2, anchor_device=None):
  """Reference utils.py:78-124: table level-1 id -> level-2 id, on every GPU."""
  devices = [c.device for c in cluster_indices_1]
  size = max(int(c.max()) for c in cluster_indices_1) + 1
  table = torch.zeros((size,), dtype=torch.long, device=devices[0])
  for c1, c2 in zip(cluster_indices_1, cluster_indices_2):
    # the reference keeps, per level-1 id, the LARGEST level-2 id (its sorted unique() writes last)
    table.scatter_reduce_(0, c1.to(devices[0]), c2.to(devices[0]), reduce='amax', include_self=True)
  return [table.to(d) for d in devices]


def gather_and_reorder_image_indices(image_indices, anchor_device=None):
  """Reference utils.py:41-74: image ids renumbered by first occurrence; every
  GPU receives the whole vector."""
  devices = [i.device for i in image_indices]
  ids = torch.cat([i.to(devices[0]) for i in image_indices], 0)
  _, inv = torch.unique(ids, return_inverse=True)
  first = torch.full((int(inv.max()) + 1,), inv.numel(), dtype=torch.long, device=devices[0])
  first.scatter_reduce_(0, inv, torch.arange(inv.numel(), device=devices[0]), reduce='amin', include_self=True)
  _, out = torch.unique(first[inv], return_inverse=True)
  return [out.to(d) for d in devices]


def gather_and_update_datas(datas, anchor_device=None):
  """Reference utils.py:220-240."""
  devices = [d.device for d in datas]
  return [torch.cat([x.to(d) for x in datas], 0) for d in devices]


# ---------------------------------------------------------------- one process per GPU
class _AllGatherVarlen(torch.autograd.Function):
  """Concatenate a [n_r, ...] tensor over ranks (n_r may differ).  Backward:
  all-reduce(sum) of the gradient, then this rank's slice -- every rank's loss
  sees every prototype."""

  @staticmethod
  def forward(ctx, t, sizes, rank, group):
    world = len(sizes)
    cap = max(sizes) if sizes else 0
    padded = t.new_zeros((cap,) + tuple(t.shape[1:]))
    padded[:t.shape[0]] = t
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    ctx.sizes, ctx.rank, ctx.group = sizes, rank, group
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)

  @staticmethod
  def backward(ctx, g):
    g = g.contiguous().clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
    start = sum(ctx.sizes[:ctx.rank])
    return g[start:start + ctx.sizes[ctx.rank]], None, None, None


def all_gather_sizes(n, device, group=None):
  world = dist.get_world_size(group)
  mine = torch.tensor([n], dtype=torch.long, device=device)
  outs = [torch.empty_like(mine) for _ in range(world)]
  dist.all_gather(outs, mine, group=group)
  return [int(o) for o in outs]


def all_gather_varlen(t, group=None, sizes=None):
  if sizes is None:
    sizes = all_gather_sizes(t.shape[0], t.device, group)
  return _AllGatherVarlen.apply(t, sizes, dist.get_rank(group), group)


class _PackedPrototypeExchange(torch.autograd.Function):
  """ONE all-gather for the whole prototype exchange: every rank sends a fixed-capacity record
  [count | prototypes | prototypes_with_loc | sem | inst | batch] (bytes), so there is no size exchange before
  the collective and no host synchronisation until the counts -- which ride in the payload -- are read to size
  the outputs.  Backward: one all-reduce of the two float gradients, then this rank's slice."""

  @staticmethod
  def forward(ctx, prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch, capacity, group):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, d = prototypes.shape
    d2 = prototypes_with_loc.shape[1]
    if n > capacity:
      raise ValueError('exchange_prototypes: %d prototypes on this rank, capacity %d' % (n, capacity))
    dev = prototypes.device
    floats = capacity * (d + d2)
    record = torch.zeros((8 + 4 * floats + 3 * 8 * capacity,), dtype=torch.uint8, device=dev)
    record[:8].view(torch.int64)[0] = n
    fl = record[8:8 + 4 * floats].view(torch.float32)
    fl[:capacity * d].view(capacity, d)[:n] = prototypes.detach().float()
    fl[capacity * d:].view(capacity, d2)[:n] = prototypes_with_loc.detach().float()
    ints = record[8 + 4 * floats:].view(torch.int64).view(3, capacity)
    ints[0, :n], ints[1, :n], ints[2, :n] = proto_sem.long(), proto_inst.long(), proto_batch.long()
    gathered = torch.empty((world * record.numel(),), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, record, group=group)
    gathered = gathered.view(world, record.numel())
    sizes = gathered[:, :8].contiguous().view(torch.int64).view(-1).tolist()          # the one host read
    fl_all = gathered[:, 8:8 + 4 * floats].contiguous().view(torch.float32).view(world, floats)
    in_all = gathered[:, 8 + 4 * floats:].contiguous().view(torch.int64).view(world, 3, capacity)
    protos = torch.cat([fl_all[r, :capacity * d].view(capacity, d)[:sizes[r]] for r in range(world)], 0)
    protos_loc = torch.cat([fl_all[r, capacity * d:].view(capacity, d2)[:sizes[r]] for r in range(world)], 0)
    labels = [torch.cat([in_all[r, j, :sizes[r]] for r in range(world)], 0) for j in range(3)]
    ctx.sizes, ctx.rank, ctx.group = sizes, rank, group
    ctx.mark_non_differentiable(*labels)
    offset = torch.tensor(sum(sizes[:rank]), dtype=torch.int64, device=dev)
    ctx.mark_non_differentiable(offset)
    return (protos, protos_loc) + tuple(labels) + (offset,)

  @staticmethod
  def backward(ctx, gp, gpl, *_unused):
    start, n = sum(ctx.sizes[:ctx.rank]), ctx.sizes[ctx.rank]
    parts = [g.contiguous().reshape(-1) for g in (gp, gpl) if g is not None]
    flat = torch.cat(parts) if parts else None
    if flat is not None:
      dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=ctx.group)
    out, at = [], 0
    for g in (gp, gpl):
      if g is None:
        out.append(None)
        continue
      out.append(flat[at:at + g.numel()].view_as(g)[start:start + n])
      at += g.numel()
    return out[0], out[1], None, None, None, None, None


def exchange_prototypes(local_ids, prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch,
                        group=None, capacity=None):
  """All-gather the per-rank prototypes and labels (rank order == the
  reference's global order) and shift this rank's pixel -> prototype ids by the
  number of prototypes on lower ranks.  Pure torch.distributed (gloo or nccl).

  capacity: a host-known upper bound on the prototypes of ANY rank, the same value on every rank (e.g. images
  per rank x slots per image).
  With it the exchange is one packed all-gather and one host read; without it, a size all-gather (with its
  host read) followed by five padded all-gathers."""
  if capacity is not None:
    res = _PackedPrototypeExchange.apply(prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch,
                                         int(capacity), group)
    return res[0], res[1], res[2], res[3], res[4], local_ids + res[5]
  sizes = all_gather_sizes(prototypes.shape[0], prototypes.device, group)
  rank = dist.get_rank(group)
  out = [all_gather_varlen(t, group, sizes)
         for t in (prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch)]
  return out[0], out[1], out[2], out[3], out[4], local_ids + sum(sizes[:rank])


class _CountedPrototypeExchange(torch.autograd.Function):
  """The prototype exchange with every count on the device: pack (one launch) -> ONE all-gather -> unpack (one
  launch).  Inputs and outputs are fixed-capacity arrays; the number of valid rows travels in the records and
  comes back as a device scalar for the NCE (ops.nce_log_likelihood(num_prototypes=...)).  No size exchange, no
  host read, no per-rank torch.cat.  Backward: one all-reduce of the two float gradients, then this rank's rows."""

  @staticmethod
  def forward(ctx, prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch, num_prototypes, capacity, group):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cap = int(capacity)
    d, d2 = prototypes.shape[1], prototypes_with_loc.shape[1]
    record = ops.exchange_pack(prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch, num_prototypes, cap)
    gathered = torch.empty((world * record.numel(),), dtype=torch.uint8, device=record.device)
    dist.all_gather_into_tensor(gathered, record, group=group)
    protos, protos_loc, sem, inst, batch, total, offset = ops.exchange_unpack(gathered, world, rank, cap, d, d2)
    ctx.group, ctx.cap, ctx.rows_in = group, cap, (prototypes.shape[0], prototypes_with_loc.shape[0])
    ctx.save_for_backward(offset, num_prototypes.reshape(-1)[:1].long())
    ctx.mark_non_differentiable(sem, inst, batch, total, offset)
    return protos, protos_loc, sem, inst, batch, total, offset

  @staticmethod
  def backward(ctx, gp, gpl, *_unused):
    offset, count = ctx.saved_tensors
    slot = torch.arange(ctx.cap, device=offset.device)
    keep = (slot < count).unsqueeze(1)
    out = []
    for g, rows_in in zip((gp, gpl), ctx.rows_in):
      if g is None:
        out.append(None)
        continue
      g = g.contiguous()
      dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
      mine = g.index_select(0, (offset + slot).clamp_(max=g.shape[0] - 1)) * keep
      if rows_in > ctx.cap:
        mine = torch.cat([mine, mine.new_zeros((rows_in - ctx.cap, mine.shape[1]))], 0)
      out.append(mine)
    return out[0], out[1], None, None, None, None, None, None


def exchange_prototypes_counted(local_ids, prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch,
                                num_prototypes, capacity, group=None):
  """exchange_prototypes for a step whose prototype count lives on the device (segment_by_kmeans_ex without
  count_prototypes, pool_prototypes at capacity).  Returns (prototypes, prototypes_with_loc, sem, inst, batch,
  shifted local ids, total) where the arrays have world*capacity rows -- the valid ones first, in rank order
  (the reference's global order, hsg/models/utils.py:172-213) -- and `total` is the device-side number of valid
  rows, to be handed to segsort_loss_multi(num_prototypes=total)."""
  res = _CountedPrototypeExchange.apply(prototypes, prototypes_with_loc, proto_sem, proto_inst, proto_batch,
                                        num_prototypes, int(capacity), group)
  return res[0], res[1], res[2], res[3], res[4], local_ids + res[6], res[5]


def dist_gather_clustering_and_update_prototypes(embeddings, embeddings_with_loc, cluster_indices,
                                                 batch_indices, semantic_labels, instance_labels,
                                                 group=None):
  """Per-rank form of gather_clustering_and_update_prototypes: tensors of THIS
  rank in, global prototypes + this rank's updated ids out."""
  ids, protos, protos_loc, psem, pinst, pbat, _ = _local_rank_and_pool(
      embeddings, embeddings_with_loc, cluster_indices, batch_indices, semantic_labels, instance_labels)
  return exchange_prototypes(ids, protos, protos_loc, psem, pinst, pbat, group)


def dist_kmeans_with_initial_labels(embeddings, initial_labels, max_label, iterations=10, group=None,
                                    collective=True):
  """Flat spherical k-means over rows sharded across ranks: each iteration all-reduces the [K,D] centroid
  sums (the one real exchange step of the path), every rank normalises identically, then assigns its own
  rows.  The sums are exact int64 fixed-point sums (ops.segment_sum_exact) and the all-reduce of integers is
  exact, so centroids and labels are bit-identical at every rank count and for every way of sharding the
  rows -- including the single-process run (`collective=False`: this process holds every row).

  Because the sums are exact integers they can also be UPDATED exactly: after the first iteration a shard only
  touches the rows whose label changed (sum[new] += x, sum[old] -= x), the all-reduce carries the integer
  corrections, and every rank adds them to its copy of the running sums -- bit-identical to re-summing every row.
  One iteration is two library calls around the collective (ops.DistKMeans / hsg_kmeans_dist_*), no host
  synchronisation."""
  x = embeddings.reshape(-1, embeddings.shape[-1]).detach()
  labels = initial_labels.reshape(-1).long()
  k = int(max_label)
  if int(iterations) <= 0:
    return labels.clone()
  d16 = ops.tc_d16(x.shape[1], k)
  xh = xerr = None
  if d16 and x.shape[0] >= 16384:
    xh, xerr = ops.make_half_copy(x, d16)
  shard = ops.DistKMeans(x, labels, k, xh=xh, xerr=xerr)
  exchange = collective and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
  running = None
  with torch.no_grad():
    for it in range(int(iterations)):
      part = shard.local()
      if exchange:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
      running = part if running is None else running.add_(part)
      shard.assign(running)
    labels = shard.labels()
  return labels
