"""Drop-in operators for hsg/utils/segsort/common.py.

Same function names, positional order, defaults, return-tuple order, dtypes
(int64 labels, float32 floats), devices and autograd-connectedness as the
reference, so ``hsg_b200.patch()`` can rebind the reference module's
attributes and pyscripts/train/train.py runs unchanged.  The arithmetic runs in
libhsgb200.so (sm_100a kernels); only index bookkeeping on tiny tensors stays
in torch.
"""

import torch

from ... import ops
from ..._lib import REDUCE_NORMALIZE, KMEANS_AUTO

_GRID_CACHE = {}
_RUN_SUMS_MIN_PIXELS = 1 << 18     # below this the k-means loop re-sums every row anyway (launch-bound regime)


def calculate_prototypes_from_labels(embeddings, labels, max_label=None):
  """Mean direction of the embeddings of each label (reference :11-41);
  doubles as the k-means M-step.  Empty label -> zero row.  Differentiable."""
  embeddings = embeddings.reshape(-1, embeddings.shape[-1])
  if max_label is None:
    max_label = labels.max() + 1                      # host sync, as in the reference
  return ops.segment_reduce(embeddings, labels.reshape(-1), int(max_label), REDUCE_NORMALIZE)


def find_nearest_prototypes(embeddings, prototypes):
  """argmax_k <x, c_k> (reference :44-64); ties -> lowest index.  int64."""
  embeddings = embeddings.reshape(-1, prototypes.shape[-1])
  with torch.no_grad():
    return ops.kmeans_estep(embeddings, prototypes.reshape(1, -1, prototypes.shape[-1]))


def kmeans_with_initial_labels(embeddings, initial_labels, max_label=None, iterations=10):
  """T x (M-step, E-step) on one [N,C] problem (reference :67-97); labels only."""
  if max_label is None:
    max_label = initial_labels.max() + 1
  max_label = int(max_label)
  with torch.no_grad():
    x = embeddings.reshape(-1, embeddings.shape[-1]).detach()
    xh = xerr = None
    d16 = ops.tc_d16(x.shape[1], max_label)
    if d16 and iterations >= 2 and x.shape[0] >= 16384:
      xh, xerr = ops.make_half_copy(x, d16)
    return ops.kmeans(x, initial_labels, max_label, int(iterations), xh=xh, xerr=xerr,
                      flags=KMEANS_AUTO)


def kmeans(embeddings, num_clusters, iterations=10):
  """Reference :100-126.  (The reference version cannot run: it calls
  initialize_cluster_labels without the device argument.)"""
  shape = embeddings.shape
  labels = initialize_cluster_labels(num_clusters, [shape[1], shape[2]], embeddings.device)
  labels = labels.view(1, shape[1], shape[2]).expand(shape[0], -1, -1).reshape(-1)
  labels = kmeans_with_initial_labels(embeddings.reshape(-1, shape[3]), labels, iterations=iterations)
  return labels.view(shape[0], shape[1], shape[2])


def initialize_cluster_labels(num_clusters, img_dimensions, device):
  """Uniform grid labels y + Ky*x (reference :129-153); computed with the same
  device-side linspace().round_() so half-way cases land where the reference's do."""
  y = torch.linspace(0, num_clusters[0] - 1, img_dimensions[0], device=device).round_().long()
  x = torch.linspace(0, num_clusters[1] - 1, img_dimensions[1], device=device).round_().long()
  return y.view(-1, 1) + (y.max() + 1) * x.view(1, -1)


def generate_location_features(img_dimensions, device, feature_type='int'):
  """[H,W,2] (y,x) coordinates (reference :156-189)."""
  if feature_type == 'int':
    y = torch.arange(img_dimensions[0], device=device)
    x = torch.arange(img_dimensions[1], device=device)
  elif feature_type == 'float':
    y = torch.linspace(0, 1, img_dimensions[0], device=device)
    x = torch.linspace(0, 1, img_dimensions[1], device=device)
  else:
    raise ValueError('Type of location features should be either int or float.')
  yy, xx = torch.meshgrid(y, x, indexing='ij')
  return torch.stack([yy, xx], dim=2)


def prepare_prototype_labels(semantic_labels, instance_labels, offset=256):
  """Reference :192-218: dense instance ids + the semantic label of each."""
  panoptic = semantic_labels + instance_labels * offset
  proto_panoptic, unique_instance = torch.unique(panoptic, return_inverse=True)
  return proto_panoptic % offset, unique_instance


def find_majority_label_index(semantic_labels, cluster_labels):
  """Reference :221-267 (inference helper, integer bookkeeping only)."""
  semantic_labels = semantic_labels.reshape(-1)
  cluster_labels = cluster_labels.reshape(-1)
  num_clusters = int(cluster_labels.max()) + 1
  num_classes = int(semantic_labels.max()) + 1
  votes = torch.zeros((num_clusters * num_classes,), dtype=torch.long, device=semantic_labels.device)
  votes.scatter_add_(0, cluster_labels * num_classes + semantic_labels, torch.ones_like(semantic_labels))
  majority = torch.argmax(votes.view(num_clusters, num_classes), 1)
  keep = torch.eq(majority[cluster_labels], semantic_labels).nonzero()
  return keep, majority


class _PrepOutputs(torch.autograd.Function):
  """Makes the two float outputs of the prep kernel differentiable in the input
  embeddings (the reference's permute/normalize/cat/normalize/index_select chain
  is; the NCE loss back-propagates through `cluster_embedding`).  Backward = one
  kernel (hsg_prep_bwd_f32)."""

  @staticmethod
  def forward(ctx, embeddings, x, xloc, pixel, loc, loc_stride, dropped):
    ctx.save_for_backward(embeddings, x, pixel, loc)
    ctx.loc_stride = loc_stride
    ctx.dropped = dropped
    return x.view_as(x), xloc.view_as(xloc)

  @staticmethod
  def backward(ctx, gx, gxloc):
    emb, y, pixel, loc = ctx.saved_tensors
    b, d, h, w = emb.shape
    row_of_pixel = None
    if ctx.dropped:                                   # inverse of the compaction: source pixel -> output row
      row_of_pixel = torch.full((b * h * w,), -1, dtype=torch.int64, device=emb.device)
      row_of_pixel[pixel] = torch.arange(pixel.shape[0], device=emb.device, dtype=torch.int64)
    gemb = ops.prep_backward(emb, y, loc, ctx.loc_stride, row_of_pixel, gx, gxloc)
    return gemb, None, None, None, None, None, None


_LOC_CACHE = {}


def _grid_init(num_clusters, hw, device):
  key = (int(num_clusters[0]), int(num_clusters[1]), int(hw[0]), int(hw[1]), str(device))
  hit = _GRID_CACHE.get(key)
  if hit is None:
    lab = initialize_cluster_labels(num_clusters, hw, device).reshape(-1)
    uniq, inv = torch.unique(lab, return_inverse=True)        # reference :341 (same for every image)
    hit = (inv.contiguous(), int(uniq.numel()))
    _GRID_CACHE[key] = hit
  return hit


def segment_by_kmeans(embeddings, labels=None, num_clusters=[5, 5], cluster_indices=None,
                      local_features=None, ignore_index=None, iterations=10):
  """Per-image spherical k-means and dense relabel (reference :270-408).

  Returns (embeddings [N,C] normalised, embeddings_with_loc [N,C+L], labels [N],
  cluster_indices [N], batch_indices [N]) with the reference's ordering
  contract: pixels image by image in raster order with ignore pixels removed;
  cluster ids are the ranks of the distinct (image, cluster, label) triples."""
  ex = segment_by_kmeans_ex(embeddings, labels, num_clusters, cluster_indices, local_features,
                            ignore_index, iterations)
  return ex['embeddings'], ex['embeddings_with_loc'], ex['labels'], ex['cluster_indices'], ex['batch_indices']


def segment_by_kmeans_ex(embeddings, labels=None, num_clusters=[5, 5], cluster_indices=None,
                         local_features=None, ignore_index=None, iterations=10, count_prototypes=False):
  """segment_by_kmeans plus the by-products the next stages can reuse (image
  offsets, per-prototype descriptions) so they need no sort / unique of their own."""
  if not embeddings.is_cuda:
    raise ops._lib.HsgError('segment_by_kmeans: CUDA tensors only (no CPU fallback)')
  b, c, h, w = embeddings.shape
  dev = embeddings.device

  if local_features is None:                                              # :313-317
    # the same read-only map for every call of a given size: seven small torch launches otherwise, a third of the
    # host time of a training-shape call
    key = (h, w, str(dev))
    loc = _LOC_CACHE.get(key)
    if loc is None:
      loc = (generate_location_features((h, w), dev, 'float') - 0.5).contiguous()
      _LOC_CACHE[key] = loc
    loc_stride = 0
  else:
    lf = local_features
    if lf.dim() == 4 and lf.stride(0) == 0:                               # expand()ed map
      loc, loc_stride = lf[0].float().contiguous(), 0
    else:
      loc = lf.float().contiguous()
      loc_stride = h * w * lf.shape[-1]
  n_loc = loc.shape[-1]

  seg_k = None
  if cluster_indices is None:                                             # :320-323, :341
    init, kmax = _grid_init(num_clusters, (h, w), dev)
    init_stride = 0
  else:                                                                   # explicit ids: per-image unique
    ci = cluster_indices.reshape(b, -1) if cluster_indices.numel() == b * h * w else \
        cluster_indices.expand(b, h, w).reshape(b, -1)
    dense, ks = [], []
    for bi in range(b):
      uniq, inv = torch.unique(ci[bi], return_inverse=True)
      dense.append(inv)
      ks.append(int(uniq.numel()))
    init = torch.stack(dense, 0).contiguous()
    init_stride = h * w
    kmax = max(ks)
    seg_k = torch.tensor(ks, dtype=torch.int32, device=dev)

  lab_in = labels.long().contiguous() if labels is not None else None
  ign = int(ignore_index) if ignore_index is not None else None           # may be a 0-dim tensor
  if lab_in is None and ign is not None:
    # the reference treats missing labels as all-zero (:325-327), so an ignore_index of 0 drops every pixel
    # and any other value drops none
    if ign == 0:
      raise RuntimeError('segment_by_kmeans: every pixel is ignored')
    ign = None
  gpu_id = dev.index or 0
  # below ~16k pixels the step is launch-bound: the fp32 CUDA-core E-step needs two launches fewer per iteration
  want_half = ops.tc_d16(c + n_loc, kmax) == c and iterations >= 1 and b * h * w >= 16384
  with torch.no_grad():
    # at full-resolution sizes the prep kernel also emits the first M-step's partial sums (the rows are on
    # chip anyway), so k-means does not start by re-reading every row it has just written
    want_runs = int(iterations) >= 1 and b * h * w >= _RUN_SUMS_MIN_PIXELS
    buf = ops.prep(embeddings.detach(), loc, loc_stride, lab_in, ign, init, init_stride,
                   b * gpu_id, want_half, want_run_sums=want_runs)        # :376-377 batch offset
    n = b * h * w if ign is None else int(buf['seg_offsets'][-1])         # one sync when pixels are dropped
    x, xloc = buf['x'][:n], buf['xloc'][:n]
    lab, bat, pix = buf['labels'][:n], buf['batch'][:n], buf['pixel'][:n]
    xh = buf['xh'][:n] if want_half else None
    xerr = buf['xerr'][:n] if want_half else None
    if n == 0:
      raise RuntimeError('segment_by_kmeans: every pixel is ignored')     # the reference fails on .max() of an empty tensor
    max_len = min(h * w, n)            # a single image with ignore pixels dropped is shorter than its grid
    clusters = ops.kmeans(xloc, buf['clusters'][:n], kmax, int(iterations),
                          seg_offsets=buf['seg_offsets'], max_seg_len=max_len, seg_k=seg_k,
                          xh=xh, xerr=xerr, flags=KMEANS_AUTO, run_sums=buf.get('runs'))
    if labels is None:
      label_values = torch.zeros((1,), dtype=torch.int64, device=dev)
    else:
      label_values = torch.unique(lab)
    ids, pl, pb, pc, npro = ops.relabel(bat, clusters, lab, b * gpu_id, b, kmax, label_values)   # :397-405

  if embeddings.requires_grad and torch.is_grad_enabled():
    x, xloc = _PrepOutputs.apply(embeddings, x, xloc, pix, loc, loc_stride, ign is not None)
  ex = {'embeddings': x, 'embeddings_with_loc': xloc, 'labels': lab, 'cluster_indices': ids,
        'batch_indices': bat, 'seg_offsets': buf['seg_offsets'], 'max_seg_len': max_len,
        'num_images': b, 'batch_base': b * gpu_id, 'kmeans_labels': clusters,
        'slots_per_image': kmax * int(label_values.numel()),
        'num_prototypes_device': npro, 'proto_label': pl, 'proto_batch': pb, 'proto_cluster': pc}
  if count_prototypes:
    p = int(npro)                                                         # the one host sync of this stage
    ex['num_prototypes'] = p
    ex['proto_label'], ex['proto_batch'], ex['proto_cluster'] = pl[:p], pb[:p], pc[:p]
  return ex


def pool_prototypes(ex, embeddings=None):
  """calculate_prototypes_from_labels(embeddings, cluster_indices) for the output
  of segment_by_kmeans_ex, using the image structure (ids are ranked by image) so
  the reduction needs no global sort.  Differentiable in the embeddings."""
  x = ex['embeddings'] if embeddings is None else embeddings
  # without the host-side count (count_prototypes=False) the result has one row per SLOT of the step
  # (ex['proto_batch'].shape[0]); rows beyond ex['num_prototypes_device'] are zero -- no host read
  num = ex['num_prototypes'] if 'num_prototypes' in ex else int(ex['proto_batch'].shape[0])
  dev = x.device
  images = torch.arange(ex['num_images'], device=dev, dtype=torch.int64) + ex['batch_base']
  seg_base = torch.searchsorted(ex['proto_batch'], images).contiguous()   # first prototype id of each image
  return ops.segment_reduce(x, ex['cluster_indices'], num, REDUCE_NORMALIZE,
                            seg_offsets=ex['seg_offsets'], max_seg_len=ex['max_seg_len'],
                            seg_base=seg_base, kmax=ex['slots_per_image'])
