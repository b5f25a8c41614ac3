"""Oracle: spherical k-means operators (numpy restatement, test infrastructure).

Restates hsg/utils/general/common.py and hsg/utils/segsort/common.py of the
reference.  All arrays are numpy; floats are float32 unless ``dtype`` says
otherwise (float64 is used by the tests to certify arg-max margins).
Label tensors are int64, like the reference.
"""

import numpy as np

EPS = 1e-12


# --------------------------------------------------------------------------
# a1  normalize_embedding            hsg/utils/general/common.py:101-120
# --------------------------------------------------------------------------
def normalize_embedding(x, eps=EPS):
  """x / max(||x||_2, eps) along the last axis (norm < eps is replaced by eps)."""
  x = np.asarray(x)
  norm = np.sqrt(np.sum(x * x, axis=-1, keepdims=True, dtype=x.dtype))
  norm = np.where(norm >= eps, norm, np.asarray(eps, dtype=x.dtype))
  return (x / norm).astype(x.dtype)


# --------------------------------------------------------------------------
# a2  location features / initial labels   hsg/utils/segsort/common.py:129-189
# --------------------------------------------------------------------------
def _linspace_f32(start, end, steps):
  """float32 linspace with torch's two-sided formula (start+i*step below the
  midpoint, end-(n-1-i)*step above it).  torch's CPU kernel vectorises the same
  formula, so values agree to 1 ulp, not bitwise."""
  if steps == 1:
    return np.asarray([start], np.float32)
  start = np.float32(start)
  end = np.float32(end)
  step = np.float32((end - start) / np.float32(steps - 1))
  i = np.arange(steps)
  lo = start + step * i.astype(np.float32)
  hi = end - step * (steps - 1 - i).astype(np.float32)
  return np.where(i < steps // 2, lo, hi).astype(np.float32)


def generate_location_features(img_dimensions, feature_type='int'):
  """[H,W,2] (y,x) coordinates: arange ('int') or linspace(0,1) ('float').
  Reference: segsort/common.py:156-189."""
  h, w = img_dimensions
  if feature_type == 'int':
    y = np.arange(h, dtype=np.int64)
    x = np.arange(w, dtype=np.int64)
  elif feature_type == 'float':
    y = _linspace_f32(0, 1, h)
    x = _linspace_f32(0, 1, w)
  else:
    raise ValueError('Type of location features should be either int or float.')
  yy, xx = np.meshgrid(y, x, indexing='ij')
  return np.stack([yy, xx], axis=2)


def initialize_cluster_labels(num_clusters, img_dimensions):
  """[H,W] int64 grid labels y + Ky*x with y,x = round_half_even(linspace).
  Reference: segsort/common.py:129-153."""
  ky, kx = num_clusters
  h, w = img_dimensions
  y = np.rint(_linspace_f32(0, ky - 1, h)).astype(np.int64).reshape(-1, 1)
  x = np.rint(_linspace_f32(0, kx - 1, w)).astype(np.int64).reshape(1, -1)
  return y + (y.max() + 1) * x


# --------------------------------------------------------------------------
# a5  calculate_prototypes_from_labels   hsg/utils/segsort/common.py:11-41
# --------------------------------------------------------------------------
def scatter_sum(x, labels, num_bins):
  """Sum rows of x into bins, rows visited in ascending order (the order of
  torch's CPU scatter_add_: checked bit-equal in gen_golden.py)."""
  x = np.asarray(x)
  x2 = x.reshape(-1, x.shape[-1])
  out = np.zeros((int(num_bins), x2.shape[1]), dtype=x2.dtype)
  np.add.at(out, np.asarray(labels).reshape(-1), x2)
  return out


def calculate_prototypes_from_labels(embeddings, labels, max_label=None):
  """normalize(scatter_sum(embeddings by label)); empty bin -> zero row."""
  labels = np.asarray(labels).reshape(-1)
  if max_label is None:
    max_label = int(labels.max()) + 1
  return normalize_embedding(scatter_sum(embeddings, labels, max_label))


# --------------------------------------------------------------------------
# a6  find_nearest_prototypes            hsg/utils/segsort/common.py:44-64
# --------------------------------------------------------------------------
def similarities(embeddings, prototypes, dtype=None):
  e = np.asarray(embeddings).reshape(-1, prototypes.shape[-1])
  p = np.asarray(prototypes)
  if dtype is not None:
    e = e.astype(dtype)
    p = p.astype(dtype)
  return e @ p.T


def find_nearest_prototypes(embeddings, prototypes, dtype=None):
  """argmax_k <x, c_k>; ties -> lowest k (np.argmax == torch.argmax rule)."""
  return np.argmax(similarities(embeddings, prototypes, dtype), axis=1).astype(np.int64)


def argmax_margins(embeddings, prototypes, chunk=1 << 16):
  """float64 certificate for an E-step: (best index, best value, gap to the
  runner-up) per pixel.  Used by the tests to decide which pixels any fp32
  implementation must agree on."""
  e = np.asarray(embeddings, np.float64).reshape(-1, prototypes.shape[-1])
  p = np.asarray(prototypes, np.float64)
  n = e.shape[0]
  best = np.empty(n, np.int64)
  val = np.empty(n, np.float64)
  gap = np.empty(n, np.float64)
  for s in range(0, n, chunk):
    sim = e[s:s + chunk] @ p.T
    b = np.argmax(sim, axis=1)
    rows = np.arange(sim.shape[0])
    v = sim[rows, b]
    if sim.shape[1] > 1:
      sim[rows, b] = -np.inf
      g = v - sim.max(axis=1)
    else:
      g = np.full_like(v, np.inf)
    best[s:s + chunk] = b
    val[s:s + chunk] = v
    gap[s:s + chunk] = g
  return best, val, gap


def similarity_to(embeddings, prototypes, labels, chunk=1 << 16):
  """float64 <x_i, c_{labels_i}> (for checking near-tie choices)."""
  e = np.asarray(embeddings, np.float64).reshape(-1, prototypes.shape[-1])
  p = np.asarray(prototypes, np.float64)
  return np.einsum('nd,nd->n', e, p[np.asarray(labels).reshape(-1)])


# --------------------------------------------------------------------------
# a4  kmeans_with_initial_labels         hsg/utils/segsort/common.py:67-97
# --------------------------------------------------------------------------
def kmeans_with_initial_labels(embeddings, initial_labels, max_label=None,
                               iterations=10, return_trace=False):
  """T x (M-step, E-step); returns labels only (like the reference)."""
  labels = np.asarray(initial_labels).reshape(-1).astype(np.int64)
  if max_label is None:
    max_label = int(labels.max()) + 1
  trace = []
  for _ in range(iterations):
    prototypes = calculate_prototypes_from_labels(embeddings, labels, max_label)
    new_labels = find_nearest_prototypes(embeddings, prototypes)
    if return_trace:
      trace.append((labels, prototypes, new_labels))
    labels = new_labels
  if return_trace:
    return labels, trace
  return labels


# --------------------------------------------------------------------------
# a7  prepare_prototype_labels           hsg/utils/segsort/common.py:192-218
# --------------------------------------------------------------------------
def prepare_prototype_labels(semantic_labels, instance_labels, offset=256):
  pan = np.asarray(semantic_labels, np.int64) + np.asarray(instance_labels, np.int64) * int(offset)
  uniq, inv = np.unique(pan, return_inverse=True)
  return (uniq % int(offset)).astype(np.int64), inv.reshape(-1).astype(np.int64)


# --------------------------------------------------------------------------
# a8  segment_mean                       hsg/utils/general/common.py:123-147
# --------------------------------------------------------------------------
def segment_mean(x, index):
  x = np.asarray(x, np.float32)
  x2 = x.reshape(-1, x.shape[-1])
  index = np.asarray(index).reshape(-1)
  m = int(index.max()) + 1
  cnt = np.zeros((m,), np.float32)
  np.add.at(cnt, index, np.float32(1))
  cnt = np.where(cnt == 0, np.float32(1), cnt)
  return (scatter_sum(x2, index, m) / cnt.reshape(-1, 1)).astype(np.float32)


# --------------------------------------------------------------------------
# a3  segment_by_kmeans                  hsg/utils/segsort/common.py:270-408
# --------------------------------------------------------------------------
def segment_by_kmeans(embeddings, labels=None, num_clusters=(5, 5),
                      cluster_indices=None, local_features=None,
                      ignore_index=None, iterations=10, gpu_id=0):
  """Per-image spherical k-means + dense global relabel.

  embeddings [B,C,H,W] float32 (NCHW).  Returns (emb [N,C], emb_with_loc
  [N,C+L], labels [N], cluster_indices [N], batch_indices [N]); pixels whose
  label == ignore_index are dropped.  ``gpu_id`` restates the reference's
  ``device.index`` batch offset (:376-377)."""
  emb = np.ascontiguousarray(np.transpose(np.asarray(embeddings, np.float32), (0, 2, 3, 1)))
  b, h, w, c = emb.shape
  emb = normalize_embedding(emb)                                     # :310
  if local_features is None:                                         # :313-317
    loc = generate_location_features((h, w), 'float').astype(np.float32) - np.float32(0.5)
    local_features = np.broadcast_to(loc.reshape(1, h, w, 2), (b, h, w, 2))
  if cluster_indices is None:                                        # :320-323
    init = initialize_cluster_labels(num_clusters, (h, w))
    cluster_indices = np.broadcast_to(init.reshape(1, h, w), (b, h, w))
  if labels is None:                                                 # :326-329
    labels = np.zeros((b, h, w), np.int64)

  out = {k: [] for k in ('lab', 'clu', 'bat', 'emb', 'loc')}
  for bi in range(b):                                                # :337
    cur_lab = np.asarray(labels[bi]).reshape(-1).astype(np.int64)
    _, cur_clu = np.unique(np.asarray(cluster_indices[bi]).reshape(-1), return_inverse=True)
    cur_clu = cur_clu.reshape(-1).astype(np.int64)
    k = int(cur_clu.max()) + 1                                       # before the ignore filter (:344)
    cur_emb = emb[bi].reshape(-1, c)
    cur_loc = np.asarray(local_features[bi], np.float32).reshape(-1, local_features.shape[-1])
    cur_xl = normalize_embedding(np.concatenate([cur_emb, cur_loc], -1))   # :349-352
    if ignore_index is not None:                                     # :355-365
      keep = np.nonzero(cur_lab != ignore_index)[0]
      cur_lab, cur_clu, cur_emb, cur_xl = cur_lab[keep], cur_clu[keep], cur_emb[keep], cur_xl[keep]
    if cur_emb.shape[0] > 0:                                         # :368
      cur_clu = kmeans_with_initial_labels(cur_xl, cur_clu, k, iterations)
    out['lab'].append(cur_lab)
    out['clu'].append(cur_clu)
    out['bat'].append(np.full_like(cur_clu, bi + b * gpu_id))        # :376-381
    out['emb'].append(cur_emb)
    out['loc'].append(cur_xl)

  lab = np.concatenate(out['lab'])
  clu = np.concatenate(out['clu'])
  bat = np.concatenate(out['bat'])
  lab_div = int(clu.max()) + 1                                       # :398
  _, clu = np.unique(bat * lab_div + clu, return_inverse=True)       # :399-401
  _, clu = prepare_prototype_labels(lab, clu.reshape(-1), int(lab.max()) + 1)   # :404-405
  return (np.concatenate(out['emb']), np.concatenate(out['loc']), lab, clu, bat)


# --------------------------------------------------------------------------
# objective used by size-independent property tests
# --------------------------------------------------------------------------
def kmeans_objective(embeddings, prototypes, labels):
  """sum_i <x_i, c_{l_i}> in float64 -- non-decreasing across E-steps."""
  return float(similarity_to(embeddings, prototypes, labels).sum())
