"""Oracle: k-NN affinity graph and DMoN loss (numpy restatement, test infrastructure).

Restates hsg/utils/graph/common.py:39-125 (affinity_matrix_as_attention) and
hsg/utils/graph/loss.py:27-88 (dmon_pool_loss) of the reference."""

import numpy as np


def exp_inner_product_kernel(x, concentration=5):
  """common.py:8-36: exp(concentration * x^T x) over the last two dims of x [B,C,n]."""
  x = np.asarray(x, np.float32)
  sim = np.einsum('bji,bjk->bik', x, x).astype(np.float32)
  return np.exp(sim * np.float32(concentration)).astype(np.float32)


def knn_adjacency(a, padding_mask=None, segment_labels=None, knn=None, remove_self_loop=True, binarize=True):
  """common.py:76-125 on a given kernel matrix a [B,n,n]."""
  a = np.array(a, np.float32, copy=True)
  b, n, _ = a.shape
  pad = np.zeros((b, n), bool) if padding_mask is None else np.asarray(padding_mask, bool)
  seg = np.zeros((b, n), np.int64) if segment_labels is None else np.asarray(segment_labels, np.int64)
  a[pad[:, :, None] | pad[:, None, :]] = 0                               # :83-85
  if remove_self_loop:                                                   # :88-97
    for i in range(b):
      if (~pad[i]).sum() > 1:
        a[i][np.eye(n, dtype=bool)] = 0
  if knn is not None:                                                    # :100-121
    for i in range(b):
      cur = a[i]
      for lab in np.unique(seg[i]):
        cols = (~pad[i]) & (seg[i] == lab)
        if not cols.any():
          continue
        k = min(int(cols.sum()), knn)
        sub = cur[:, cols]
        kth = -np.sort(-sub, axis=1)[:, k - 1]
        cur[np.ix_(np.ones(n, bool), cols)] = np.where(sub < kth[:, None], 0, sub)
      a[i] = cur
  if binarize:                                                           # :123-125
    a = (a > 0).astype(np.float32)
  return a


def dmon_pool_loss(adj, s, mask):
  """loss.py:27-88 with s already a soft assignment [B,n,k] and mask [B,n] (valid nodes)."""
  adj = np.asarray(adj, np.float64)
  s = np.asarray(s, np.float64) * np.asarray(mask, np.float64)[:, :, None]
  b, n, k = s.shape
  out_adj = np.einsum('bik,bij,bjl->bkl', s, adj, s)
  d = adj.sum(2)
  sd = np.einsum('bik,bi->bk', s, d)
  norm = 2 * d.sum(1)
  numer = np.trace(out_adj, axis1=1, axis2=2) - (sd ** 2).sum(1) / norm
  dmon = np.mean(1 - numer / norm)
  collapse = np.mean(np.linalg.norm(s.sum(1), axis=1) / (n / np.sqrt(k)))
  return dmon, collapse
