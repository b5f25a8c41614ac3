"""Torch-facing wrappers of the C ABI (device memory, streams and autograd glue).

Everything here hands raw device pointers of contiguous CUDA tensors to
libhsgb200.so on the current stream.  There is no CPU path: CPU tensors raise.
"""

import ctypes

import torch

from . import _lib
from ._lib import check


XH_TAIL = 16      # HSG_XH_TAIL in include/hsg_b200.h


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
  # the raw handle of torch's current stream (torch.cuda.current_stream() builds a Stream object: 15 us per call,
  # a fifth of the host time of a training-shape segment_by_kmeans)
  return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _need_cuda(*tensors):
  for t in tensors:
    if t is not None and not t.is_cuda:
      raise _lib.HsgError('hsg_b200 operators run on CUDA tensors only (got a %s tensor); '
                          'there is no CPU fallback' % t.device.type)


def _f32(t):
  return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _i64(t):
  return t.contiguous() if t.dtype == torch.int64 else t.long().contiguous()


def _workspace(nbytes, device):
  return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------- a1 normalize
class _Normalize(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x):
    x2 = _f32(x).view(-1, x.shape[-1])
    y = torch.empty_like(x2)
    with torch.cuda.device(x.device):
      check(_lib.load().hsg_normalize_f32(_ptr(x2), _ptr(y), x2.shape[0], x2.shape[1], _stream()),
            'normalize')
    ctx.save_for_backward(x2)
    ctx.shape = x.shape
    return y.view(x.shape)

  @staticmethod
  def backward(ctx, gy):
    (x2,) = ctx.saved_tensors
    g2 = _f32(gy).view(-1, x2.shape[1])
    gx = torch.empty_like(x2)
    with torch.cuda.device(x2.device):
      check(_lib.load().hsg_normalize_bwd_f32(_ptr(x2), _ptr(g2), _ptr(gx), x2.shape[0], x2.shape[1],
                                              _stream()), 'normalize_bwd')
    return gx.view(ctx.shape)


def normalize(x):
  _need_cuda(x)
  if x.numel() == 0:
    return x.float()
  return _Normalize.apply(x)


# ---------------------------------------------------------------- K0 prep
def prep(embeddings, loc, loc_image_stride, labels, ignore_index, init_clusters, init_image_stride,
         batch_index_base, want_half, want_run_sums=False):
  """Front half of segment_by_kmeans.  Returns a dict of max-size buffers plus
  the device offsets; the caller slices with N = seg_offsets[-1].  With want_run_sums the kernel
  also emits the partial sums of the first k-means M-step (out['runs'], for kmeans(run_sums=...))."""
  _need_cuda(embeddings, loc, labels, init_clusters)
  emb = _f32(embeddings)
  b, d, h, w = emb.shape
  l = loc.shape[-1]
  dev = emb.device
  n_max = b * h * w
  out = {
      'x': torch.empty((n_max, d), dtype=torch.float32, device=dev),
      'xloc': torch.empty((n_max, d + l), dtype=torch.float32, device=dev),
      'labels': torch.empty((n_max,), dtype=torch.int64, device=dev),
      'clusters': torch.empty((n_max,), dtype=torch.int64, device=dev),
      'batch': torch.empty((n_max,), dtype=torch.int64, device=dev),
      'pixel': torch.empty((n_max,), dtype=torch.int64, device=dev),
      'seg_offsets': torch.empty((b + 1,), dtype=torch.int64, device=dev),
      'xh': torch.empty((n_max, d + XH_TAIL), dtype=torch.float16, device=dev) if want_half else None,
      'xerr': torch.empty((n_max,), dtype=torch.float32, device=dev) if want_half else None,
  }
  lib = _lib.load()
  ws = _workspace(lib.hsg_prep_workspace_bytes(b, h, w), dev)
  use_ignore = ignore_index is not None
  args = (_ptr(emb), b, d, h, w, _ptr(loc), l, int(loc_image_stride),
          _ptr(labels), int(use_ignore), int(ignore_index) if use_ignore else 0,
          _ptr(init_clusters), int(init_image_stride), int(batch_index_base),
          _ptr(out['x']), _ptr(out['xloc']), _ptr(out['xh']), _ptr(out['xerr']),
          _ptr(out['labels']), _ptr(out['clusters']), _ptr(out['batch']), _ptr(out['pixel']), _ptr(out['seg_offsets']),
          _ptr(ws), ws.numel())
  with torch.cuda.device(dev):
    if want_run_sums:
      rpi = int(lib.hsg_prep_runs_per_image(h, w))
      runs = {'sums': torch.empty((b * rpi, d + l), dtype=torch.float32, device=dev),
              'cluster': torch.empty((b * rpi,), dtype=torch.int32, device=dev),
              'count': torch.empty((b * rpi,), dtype=torch.int32, device=dev),
              'overflow': torch.empty((1,), dtype=torch.int32, device=dev), 'per_segment': rpi}
      check(lib.hsg_prep_sums_f32(*args, _ptr(runs['sums']), _ptr(runs['cluster']), _ptr(runs['count']),
                                  _ptr(runs['overflow']), _stream()), 'prep_sums')
      out['runs'] = runs
    else:
      check(lib.hsg_prep_f32(*args, _stream()), 'prep')
  return out


def prep_backward(embeddings, x, loc, loc_image_stride, row_of_pixel, grad_x, grad_xloc):
  """Gradient of the prep chain w.r.t. the NCHW embeddings (hsg_prep_bwd_f32); either gradient may be None."""
  _need_cuda(embeddings, x, loc, row_of_pixel, grad_x, grad_xloc)
  emb = _f32(embeddings)
  b, d, h, w = emb.shape
  l = loc.shape[-1]
  gemb = torch.empty_like(emb)
  gx = _f32(grad_x) if grad_x is not None else None
  gz = _f32(grad_xloc) if grad_xloc is not None else None
  with torch.cuda.device(emb.device):
    check(_lib.load().hsg_prep_bwd_f32(_ptr(emb), b, d, h, w, _ptr(_f32(x)), _ptr(loc), l, int(loc_image_stride),
                                       _ptr(row_of_pixel), _ptr(gx), _ptr(gz), _ptr(gemb), _stream()), 'prep_bwd')
  return gemb


def make_half_copy(x, d16):
  _need_cuda(x)
  x = _f32(x)
  xh = torch.empty((x.shape[0], d16 + XH_TAIL), dtype=torch.float16, device=x.device)
  xerr = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
  with torch.cuda.device(x.device):
    check(_lib.load().hsg_make_half_copy_f32(_ptr(x), x.shape[0], x.shape[1], d16, _ptr(xh),
                                             _ptr(xerr), _stream()), 'half_copy')
  return xh, xerr


# ---------------------------------------------------------------- K1 k-means
def _seg_args(n, seg_offsets, max_seg_len, device):
  if seg_offsets is None:
    seg_offsets = torch.tensor([0, n], dtype=torch.int64, device=device)
    max_seg_len = n
  return seg_offsets, seg_offsets.numel() - 1, int(max_seg_len)


def tc_d16(dim, kmax):
  """Width of the fp16 side copy the tensor-core E-step wants for this shape (0 = none)."""
  for d16 in (512, 256, 128, 64):
    if dim >= d16 and dim - d16 <= 5 and kmax <= 255 * (256 if d16 <= 256 else 128):
      return d16
  return 0


def kmeans(x, init_labels, kmax, iterations, seg_offsets=None, max_seg_len=None, seg_k=None,
           xh=None, xerr=None, flags=_lib.KMEANS_AUTO, return_centroids=False, run_sums=None):
  _need_cuda(x, init_labels, seg_offsets, seg_k, xh, xerr)
  x = _f32(x)
  n, dim = x.shape
  init_labels = _i64(init_labels).view(-1)
  seg_offsets, s, max_seg_len = _seg_args(n, seg_offsets, max_seg_len, x.device)
  labels = torch.empty((n,), dtype=torch.int64, device=x.device)
  cent = torch.empty((s, kmax, dim), dtype=torch.float32, device=x.device) if return_centroids else None
  if n == 0 or iterations == 0:
    labels.copy_(init_labels)
    return (labels, cent) if return_centroids else labels
  lib = _lib.load()
  ws = _workspace(lib.hsg_kmeans_workspace_bytes(n, dim, s, kmax, max_seg_len), x.device)
  d16 = xh.shape[1] - XH_TAIL if xh is not None else 0
  args = (_ptr(x), n, dim, _ptr(xh), d16, _ptr(xerr), _ptr(seg_offsets), s, max_seg_len, _ptr(seg_k), kmax,
          _ptr(init_labels), iterations, _ptr(labels), _ptr(cent), flags, _ptr(ws), ws.numel())
  with torch.cuda.device(x.device):
    if run_sums is not None:
      r = run_sums
      check(lib.hsg_kmeans_presummed_f32(*args, _ptr(r['sums']), _ptr(r['cluster']), _ptr(r['count']),
                                         _ptr(r['overflow']), int(r['per_segment']), _stream()), 'kmeans_presummed')
    else:
      check(lib.hsg_kmeans_f32(*args, _stream()), 'kmeans')
  return (labels, cent) if return_centroids else labels


def kmeans_mstep(x, labels, kmax, seg_offsets=None, max_seg_len=None):
  _need_cuda(x, labels, seg_offsets)
  x = _f32(x)
  n, dim = x.shape
  labels = _i64(labels).view(-1)
  seg_offsets, s, max_seg_len = _seg_args(n, seg_offsets, max_seg_len, x.device)
  cent = torch.empty((s, kmax, dim), dtype=torch.float32, device=x.device)
  lib = _lib.load()
  ws = _workspace(lib.hsg_segment_reduce_workspace_bytes(n, dim, s * kmax, s, kmax, max_seg_len), x.device)
  with torch.cuda.device(x.device):
    check(lib.hsg_kmeans_mstep_f32(_ptr(x), n, dim, _ptr(seg_offsets), s, max_seg_len, None, kmax,
                                   _ptr(labels), _ptr(cent), _ptr(ws), ws.numel(), _stream()), 'mstep')
  return cent


def kmeans_estep(x, centroids, seg_offsets=None, max_seg_len=None, seg_k=None, xh=None, xerr=None,
                 flags=_lib.KMEANS_AUTO, return_rechecked=False):
  _need_cuda(x, centroids, seg_offsets, seg_k, xh, xerr)
  x = _f32(x)
  n, dim = x.shape
  centroids = _f32(centroids)
  seg_offsets, s, max_seg_len = _seg_args(n, seg_offsets, max_seg_len, x.device)
  kmax = centroids.shape[-2]
  assert centroids.numel() == s * kmax * dim
  labels = torch.empty((n,), dtype=torch.int64, device=x.device)
  nre = torch.zeros((2,), dtype=torch.int64, device=x.device)
  lib = _lib.load()
  ws = _workspace(lib.hsg_kmeans_workspace_bytes(n, dim, s, kmax, max_seg_len), x.device)
  d16 = xh.shape[1] - XH_TAIL if xh is not None else 0
  with torch.cuda.device(x.device):
    check(lib.hsg_kmeans_estep_f32(_ptr(x), n, dim, _ptr(xh), d16, _ptr(xerr), _ptr(seg_offsets), s,
                                   max_seg_len, _ptr(seg_k), kmax, _ptr(centroids), _ptr(labels),
                                   _ptr(nre), flags, _ptr(ws), ws.numel(), _stream()), 'estep')
  return (labels, nre) if return_rechecked else labels


# ---------------------------------------------------------------- K3 segmented reduction
class _SegmentReduce(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, labels, num_bins, mode, seg_offsets, max_seg_len, seg_base, kmax):
    x2 = _f32(x)
    n, dim = x2.shape
    dev = x2.device
    if seg_offsets is None:
      seg_offsets = torch.tensor([0, n], dtype=torch.int64, device=dev)
      seg_base = torch.zeros((1,), dtype=torch.int64, device=dev)
      max_seg_len, kmax = n, num_bins
    s = seg_offsets.numel() - 1
    out = torch.empty((num_bins, dim), dtype=torch.float32, device=dev)
    sums = torch.empty_like(out) if mode == _lib.REDUCE_NORMALIZE else None
    counts = torch.empty((num_bins,), dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = _workspace(lib.hsg_segment_reduce_workspace_bytes(n, dim, num_bins, s, kmax, max_seg_len), dev)
    with torch.cuda.device(dev):
      check(lib.hsg_segment_reduce_f32(_ptr(x2), n, dim, _ptr(labels), num_bins, _ptr(seg_offsets), s,
                                       int(max_seg_len), _ptr(seg_base), int(kmax), mode, _ptr(out),
                                       _ptr(sums), _ptr(counts), _ptr(ws), ws.numel(), _stream()),
            'segment_reduce')
    ctx.save_for_backward(out, sums, counts, labels)
    ctx.mode = mode
    ctx.n = n
    return out

  @staticmethod
  def backward(ctx, g):
    out, sums, counts, labels = ctx.saved_tensors
    p, dim = out.shape
    g = _f32(g)
    gx = torch.empty((ctx.n, dim), dtype=torch.float32, device=out.device)
    ws = _workspace(p * dim * 4, out.device)
    with torch.cuda.device(out.device):
      check(_lib.load().hsg_segment_reduce_bwd_f32(_ptr(g), _ptr(out), _ptr(sums), _ptr(counts),
                                                   _ptr(labels), ctx.n, dim, p, ctx.mode, _ptr(gx),
                                                   _ptr(ws), ws.numel(), _stream()), 'segment_reduce_bwd')
    return gx, None, None, None, None, None, None, None


MAX_BINS_PER_SEGMENT = 49152      # SR_MAX_KEYS (csrc/segreduce.cuh): shared-memory histogram of one segment's keys


def segment_reduce(x, labels, num_bins, mode, seg_offsets=None, max_seg_len=None, seg_base=None, kmax=None):
  """out[k] = finish(sum of rows of x with label k); differentiable in x."""
  _need_cuda(x, labels, seg_offsets, seg_base)
  x2 = x.reshape(-1, x.shape[-1])
  labels = _i64(labels).reshape(-1)
  if x2.shape[0] == 0:
    return torch.zeros((num_bins, x2.shape[1]), dtype=torch.float32, device=x.device)
  if seg_offsets is None and num_bins > MAX_BINS_PER_SEGMENT:
    # The reference takes any number of labels (calculate_prototypes_from_labels over the dense prototype ids of a
    # whole batch: 48 x 256 x regions).  The kernels histogram at most 49152 bins per segment, so consecutive blocks
    # of 49152 label ids become the segments.  Ids made by segment_by_kmeans are ranked by image -- the rows are already
    # grouped by block and nothing moves; otherwise the rows are gathered into block order first (sums do not depend on it).
    block = labels // MAX_BINS_PER_SEGMENT
    n_blocks = (int(num_bins) + MAX_BINS_PER_SEGMENT - 1) // MAX_BINS_PER_SEGMENT
    if not bool((block[1:] >= block[:-1]).all()):
      perm = torch.argsort(block, stable=True)
      x2, labels, block = x2.index_select(0, perm), labels.index_select(0, perm), block.index_select(0, perm)
    counts = torch.bincount(block, minlength=n_blocks)
    seg_offsets = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)])
    seg_base = torch.arange(n_blocks, dtype=torch.int64, device=x.device) * MAX_BINS_PER_SEGMENT
    max_seg_len, kmax = int(counts.max()), MAX_BINS_PER_SEGMENT
  elif (kmax if kmax is not None else num_bins) > MAX_BINS_PER_SEGMENT:
    raise _lib.HsgError('segment_reduce: more than 49152 bins per segment')
  return _SegmentReduce.apply(x2, labels, int(num_bins), mode, seg_offsets, max_seg_len, seg_base, kmax)


FIXED_POINT_SCALE = 2.0 ** -36


class DistKMeans:
  """One shard of the row-sharded flat k-means, an iteration at a time (hsg_kmeans_dist_*): `local()` returns this
  shard's exact int64 contribution to the [K,D] centroid sums (the full sums first, afterwards only the rows whose
  label changed), the caller all-reduces it and adds it to its running sums, `assign(running)` runs the E-step.
  The loop's state (labels, previous labels, sort tiles) lives in the workspace this object owns."""

  def __init__(self, x, init_labels, kmax, xh=None, xerr=None, flags=_lib.KMEANS_AUTO):
    _need_cuda(x, init_labels, xh, xerr)
    self.x = _f32(x)
    self.n, self.dim = self.x.shape
    self.k = int(kmax)
    self.init = _i64(init_labels).view(-1)
    self.xh, self.xerr, self.flags = xh, xerr, flags
    self.d16 = xh.shape[1] - XH_TAIL if xh is not None else 0
    self.off = torch.tensor([0, self.n], dtype=torch.int64, device=self.x.device)
    self.lib = _lib.load()
    self.ws = _workspace(self.lib.hsg_kmeans_dist_workspace_bytes(self.n, self.dim, self.k), self.x.device)
    self.first = True

  def local(self):
    out = torch.empty((self.k, self.dim), dtype=torch.int64, device=self.x.device)
    with torch.cuda.device(self.x.device):
      check(self.lib.hsg_kmeans_dist_local_i64(_ptr(self.x), self.n, self.dim, self.d16, _ptr(self.off), self.k,
                                               int(self.first), _ptr(self.init), _ptr(out), _ptr(self.ws),
                                               self.ws.numel(), _stream()), 'kmeans_dist_local')
    self.first = False
    return out

  def assign(self, running_sums):
    sums = running_sums.contiguous()
    assert sums.dtype == torch.int64 and sums.numel() == self.k * self.dim
    with torch.cuda.device(self.x.device):
      check(self.lib.hsg_kmeans_dist_assign_f32(_ptr(self.x), self.n, self.dim, _ptr(self.xh), self.d16, _ptr(self.xerr),
                                                _ptr(self.off), self.k, _ptr(sums), self.flags, _ptr(self.ws),
                                                self.ws.numel(), _stream()), 'kmeans_dist_assign')

  def labels(self):
    out = torch.empty((self.n,), dtype=torch.int64, device=self.x.device)
    with torch.cuda.device(self.x.device):
      check(self.lib.hsg_kmeans_dist_labels_i64(self.n, self.dim, self.d16, self.k, _ptr(out), _ptr(self.ws),
                                                self.ws.numel(), _stream()), 'kmeans_dist_labels')
    return out


def segment_sum_exact(x, labels, num_bins):
  """Exact bin sums as int64 fixed point (value = sum * FIXED_POINT_SCALE): independent of the row order and
  of how the rows are split over calls or GPUs (hsg_segment_sum_exact_i64).  x rows with |x| <= 1."""
  _need_cuda(x, labels)
  x2 = _f32(x.reshape(-1, x.shape[-1]).detach())
  labels = _i64(labels).reshape(-1)
  n, dim = x2.shape
  dev = x2.device
  if num_bins > 49152:
    raise _lib.HsgError('segment_sum_exact: more than 49152 bins')
  out = torch.empty((num_bins, dim), dtype=torch.int64, device=dev)
  seg_offsets = torch.tensor([0, n], dtype=torch.int64, device=dev)
  seg_base = torch.zeros((1,), dtype=torch.int64, device=dev)
  lib = _lib.load()
  ws = _workspace(lib.hsg_segment_sum_exact_workspace_bytes(n, dim, num_bins, 1, num_bins, n), dev)
  with torch.cuda.device(dev):
    check(lib.hsg_segment_sum_exact_i64(_ptr(x2), n, dim, _ptr(labels), num_bins, _ptr(seg_offsets), 1, n,
                                        _ptr(seg_base), int(num_bins), _ptr(out), _ptr(ws), ws.numel(), _stream()),
          'segment_sum_exact')
  return out


# ---------------------------------------------------------------- K4 NCE
class _Nce(torch.autograd.Function):

  @staticmethod
  def forward(ctx, e, protos, inst, sem, psem, plus, conc, count=None):
    e2 = _f32(e)
    p2 = _f32(protos)
    n, dim = e2.shape
    p = p2.shape[0]
    n_sets = sem.shape[0]
    dev = e2.device
    out = torch.empty((n_sets, n), dtype=torch.float32, device=dev)
    stats = torch.empty((n_sets, n, 4), dtype=torch.float32, device=dev)
    plus_arr = (ctypes.c_int32 * n_sets)(*[int(v) for v in plus])
    lib = _lib.load()
    ws = _workspace(lib.hsg_nce_workspace_bytes(n, p, dim, n_sets), dev)
    with torch.cuda.device(dev):
      check(lib.hsg_nce_fwd_counted_f32(_ptr(e2), _ptr(p2), n, p, _ptr(count) if count is not None else None, dim,
                                        _ptr(inst), _ptr(sem), _ptr(psem),
                                        n_sets, plus_arr, float(conc), _ptr(out), _ptr(stats), _ptr(ws), ws.numel(),
                                        _stream()), 'nce_fwd')
    ctx.save_for_backward(e2, p2, inst, sem, psem, stats)
    ctx.count = count
    ctx.plus = [int(v) for v in plus]
    ctx.conc = float(conc)
    return out

  @staticmethod
  def backward(ctx, g):
    e2, p2, inst, sem, psem, stats = ctx.saved_tensors
    n, dim = e2.shape
    p = p2.shape[0]
    n_sets = sem.shape[0]
    w = _f32(g)
    ge = torch.empty_like(e2)
    gp = torch.empty_like(p2)
    lib = _lib.load()
    ws = _workspace(lib.hsg_nce_workspace_bytes(n, p, dim, n_sets), e2.device)
    plus_arr = (ctypes.c_int32 * n_sets)(*ctx.plus)
    with torch.cuda.device(e2.device):
      check(lib.hsg_nce_bwd_counted_f32(_ptr(e2), _ptr(p2), n, p, _ptr(ctx.count) if ctx.count is not None else None,
                                        dim, _ptr(inst), _ptr(sem), _ptr(psem), n_sets,
                                        plus_arr, ctx.conc, _ptr(stats), _ptr(w), _ptr(ge), _ptr(gp), _ptr(ws),
                                        ws.numel(), _stream()), 'nce_bwd')
    return ge, gp, None, None, None, None, None, None


def nce_log_likelihood(embeddings, instance_labels, semantic_label_sets, prototypes,
                       prototype_semantic_label_sets, concentration, group_modes, num_prototypes=None):
  """Per-pixel negative log-likelihood for several label sets in one pass.

  semantic_label_sets [n_sets,N], prototype_semantic_label_sets [n_sets,P];
  returns [n_sets,N] float32, differentiable in embeddings and prototypes.
  num_prototypes: optional int64 device tensor [1] -- only the first num_prototypes rows of `prototypes` (and of the
  label sets) exist; the rest is capacity (no host read of the count: hsg_nce_fwd_counted_f32)."""
  _need_cuda(embeddings, prototypes, instance_labels, semantic_label_sets, prototype_semantic_label_sets)
  e = embeddings.reshape(-1, embeddings.shape[-1])
  p = prototypes.reshape(-1, prototypes.shape[-1])
  inst = _i64(instance_labels).reshape(-1)
  sem = _i64(semantic_label_sets).reshape(-1, e.shape[0])
  psem = _i64(prototype_semantic_label_sets).reshape(-1, p.shape[0])
  plus = [1 if m == 'segsort+' else 0 for m in group_modes]
  assert sem.shape[0] == psem.shape[0] == len(plus)
  if e.shape[0] == 0:
    return torch.zeros((sem.shape[0], 0), dtype=torch.float32, device=e.device)
  count = None
  if num_prototypes is not None:
    count = _i64(num_prototypes).reshape(-1)[:1]
    _need_cuda(count)
  return _Nce.apply(e, p, inst, sem, psem, plus, concentration, count)


# ---------------------------------------------------------------- K2 relabel
def relabel(batch, cluster, label, batch_base, num_images, kmax, label_values):
  """ids = rank of (batch, cluster, label) among the distinct triples present.
  Returns (ids [N], proto_label, proto_batch, proto_cluster, n_protos_device);
  the proto_* buffers are max-size, valid up to n_protos."""
  _need_cuda(batch, cluster, label, label_values)
  dev = batch.device
  n = batch.numel()
  nl = label_values.numel()
  t = num_images * kmax * nl
  cap = min(t, max(n, 1))
  ids = torch.empty((n,), dtype=torch.int64, device=dev)
  # slots beyond n_protos keep the fill: proto_batch stays sorted (searchsorted over the whole buffer is valid
  # without knowing the count on the host), labels / clusters match nothing
  # (one fill for the three descriptors: a large positive value)
  desc = torch.full((3, cap), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
  pl, pb, pc = desc[0], desc[1], desc[2]
  npro = torch.empty((1,), dtype=torch.int64, device=dev)
  lib = _lib.load()
  ws = _workspace(lib.hsg_relabel_workspace_bytes(num_images, kmax, nl), dev)
  with torch.cuda.device(dev):
    check(lib.hsg_relabel_i64(_ptr(batch), _ptr(cluster), _ptr(label), n, int(batch_base), num_images,
                              kmax, _ptr(label_values), nl, _ptr(ids), _ptr(pl), _ptr(pb), _ptr(pc),
                              _ptr(npro), _ptr(ws), ws.numel(), _stream()), 'relabel')
  return ids, pl, pb, pc, npro


# ---------------------------------------------------------------- a13 prototype exchange records
def exchange_record_bytes(capacity, dim, dim_loc):
  return int(_lib.load().hsg_exchange_record_bytes(int(capacity), int(dim), int(dim_loc)))


def exchange_pack(prototypes, prototypes_with_loc, sem, inst, batch, num_prototypes, capacity):
  """This rank's fixed-capacity record (uint8 tensor) for the one all-gather of the prototype exchange; the count
  is read on the device (include/hsg_b200.h: hsg_exchange_pack)."""
  _need_cuda(prototypes, prototypes_with_loc, sem, inst, batch, num_prototypes)
  p, pl = _f32(prototypes.detach()), _f32(prototypes_with_loc.detach())
  cap = int(capacity)
  if min(p.shape[0], pl.shape[0], sem.numel(), inst.numel(), batch.numel()) < cap:
    raise ValueError('exchange_pack: inputs hold fewer than capacity=%d rows' % cap)
  rec = torch.empty((exchange_record_bytes(cap, p.shape[1], pl.shape[1]),), dtype=torch.uint8, device=p.device)
  with torch.cuda.device(p.device):
    check(_lib.load().hsg_exchange_pack(_ptr(p), _ptr(pl), _ptr(_i64(sem)), _ptr(_i64(inst)), _ptr(_i64(batch)),
                                        _ptr(_i64(num_prototypes)), cap, p.shape[1], pl.shape[1], _ptr(rec), _stream()),
          'exchange_pack')
  return rec


def exchange_unpack(gathered, world, rank, capacity, dim, dim_loc):
  """(prototypes, prototypes_with_loc, sem, inst, batch, total [1], offset [1]) from the `world` gathered records:
  world*capacity rows, valid ones first in rank order (hsg_exchange_unpack)."""
  _need_cuda(gathered)
  dev = gathered.device
  rows = int(world) * int(capacity)
  protos = torch.empty((rows, dim), dtype=torch.float32, device=dev)
  protos_loc = torch.empty((rows, dim_loc), dtype=torch.float32, device=dev)
  sem, inst, batch = (torch.empty((rows,), dtype=torch.int64, device=dev) for _ in range(3))
  total, offset = (torch.empty((1,), dtype=torch.int64, device=dev) for _ in range(2))
  with torch.cuda.device(dev):
    check(_lib.load().hsg_exchange_unpack(_ptr(gathered), int(world), int(rank), int(capacity), int(dim), int(dim_loc),
                                          _ptr(protos), _ptr(protos_loc), _ptr(sem), _ptr(inst), _ptr(batch),
                                          _ptr(total), _ptr(offset), _stream()), 'exchange_unpack')
  return protos, protos_loc, sem, inst, batch, total, offset


# ---------------------------------------------------------------- f3 top-k retrieval
def topk_affinity(embeddings, prototypes, k):
  """Indices [N,k] (int64, best first) of the k prototypes with the largest inner product per row."""
  _need_cuda(embeddings, prototypes)
  e = _f32(embeddings.detach().reshape(-1, embeddings.shape[-1]))
  p = _f32(prototypes.detach().reshape(-1, prototypes.shape[-1]))
  idx = torch.empty((e.shape[0], int(k)), dtype=torch.int64, device=e.device)
  with torch.cuda.device(e.device):
    check(_lib.load().hsg_topk_affinity_f32(_ptr(e), e.shape[0], _ptr(p), p.shape[0], e.shape[1], int(k), _ptr(idx), None,
                                            _stream()), 'topk')
  return idx
