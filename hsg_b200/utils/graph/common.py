"""Drop-in for hsg/utils/graph/common.py: the k-NN affinity graph of the DMoN regulariser.

`affinity_matrix_as_attention` keeps the reference's signature (:39-125).  The kernel matrix
itself (`kernel_fn`, by default exp(5 x^T x): one small batched matmul) is evaluated as in the
reference; the masking / per-segment top-k / binarise part -- a Python double loop over batch
entries and segments with two host synchronisations per segment in the reference -- is one
launch of `hsg_knn_adjacency_f32`.  The result carries no gradient, like the reference's
binarised graph (`torch.where(A > 0, ones, zeros)`).
"""

import torch

from ... import _lib
from ...ops import _need_cuda, _ptr, _stream, check


def inner_product_kernel(x):
  """sim(i, j) = x_i^T x_j over the last two dimensions (reference :8-20)."""
  return torch.einsum('...ij,...jk->...ik', x.transpose(-2, -1), x)


def exp_inner_product_kernel(x, concentration=5):
  """sim(i, j) = exp(concentration * x_i^T x_j) (reference :23-36)."""
  return inner_product_kernel(x).mul(concentration).exp()


def knn_adjacency(affinity, x_padding_mask=None, x_segment_labels=None, knn=None, remove_self_loop=True,
                  binarize=True):
  """[B,n,n] kernel matrix -> masked / k-NN-per-segment / binarised graph (one launch)."""
  _need_cuda(affinity, x_padding_mask, x_segment_labels)
  a = affinity.detach().float().contiguous()
  b, n, n2 = a.shape
  if n != n2:
    raise ValueError('knn_adjacency: affinity must be [B,n,n]')
  pad = x_padding_mask.to(torch.uint8).contiguous() if x_padding_mask is not None else None
  seg = x_segment_labels.long().contiguous() if x_segment_labels is not None else None
  out = torch.empty_like(a)
  with torch.cuda.device(a.device):
    check(_lib.load().hsg_knn_adjacency_f32(_ptr(a), _ptr(pad), _ptr(seg), b, n, int(knn) if knn is not None else 0,
                                            1 if remove_self_loop else 0, 1 if binarize else 0, _ptr(out), _stream()),
          'knn_adjacency')
  return out


def affinity_matrix_as_attention(x, x_padding_mask=None, x_segment_labels=None, knn=None, remove_self_loop=True,
                                 binarize=True, kernel_fn=exp_inner_product_kernel):
  """Reference :39-125; x is [batch_size, channels, num_nodes]."""
  a = kernel_fn(x)
  if not binarize and a.requires_grad:
    # un-binarised graphs keep their values (and gradients) in the reference: select, do not replace
    keep = knn_adjacency(a, x_padding_mask, x_segment_labels, knn, remove_self_loop, False) != 0
    return torch.where(keep, a, torch.zeros_like(a))
  return knn_adjacency(a, x_padding_mask, x_segment_labels, knn, remove_self_loop, binarize)
