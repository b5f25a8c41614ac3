"""Drop-in for hsg/utils/graph/loss.py: DMoN clustering regulariser (reference :27-145).

The graph comes from the one-launch k-NN kernel (`common.affinity_matrix_as_attention`); the
pooling loss itself is a handful of [B,n,k]-sized contractions (k = 4..8 clusters) evaluated with
torch so autograd carries the gradient to the assignment logits, as in the reference.
"""

import torch
from torch.nn.modules.loss import _Loss

from . import common as graph_common


def dmon_pool_loss(x, adj, s, mask=None, softmax=False):
  """(dmon_loss, collapse_loss) of soft assignments s [B,n,k] on graph adj [B,n,n]; reference :27-88.
  `x` is accepted for signature compatibility (the reference does not use it either)."""
  adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
  s = s.unsqueeze(0) if s.dim() == 2 else s
  batch_size, num_nodes, k = s.shape
  if softmax:
    s = torch.softmax(s, dim=-1)
  if mask is not None:
    s = s * mask.view(batch_size, num_nodes, 1).to(s.dtype)
  out_adj = torch.matmul(torch.matmul(s.transpose(1, 2), adj), s)          # C^T A C
  d_flat = adj.sum(dim=2)
  sd = torch.einsum('bik,bi->bk', s, d_flat)                              # C^T d  (C^T d d^T C = its outer product)
  normalizer = 2 * d_flat.sum(dim=1)
  numerator = torch.einsum('ijj->i', out_adj) - (sd * sd).sum(dim=1) / normalizer
  dmon_loss = torch.mean(1 - numerator / normalizer)
  collapse_loss = torch.mean(torch.norm(s.sum(dim=1), dim=1) / (num_nodes / k ** 0.5))
  return dmon_loss, collapse_loss


class DMonLoss(_Loss):

  def __init__(self, adj_knn=None, size_average=None, reduce=None, reduction='mean'):
    super(DMonLoss, self).__init__(size_average, reduce, reduction)
    self._knn = adj_knn

  def __repr__(self):
    return 'DMonLoss(adj_knn={})'.format(self._knn)

  def forward(self, logits, x, x_padding_mask=None, x_segment_labels=None):
    """logits [B,k,n] soft assignments, x [B,C,n] node features; reference :116-145."""
    adj = graph_common.affinity_matrix_as_attention(
        x, x_padding_mask, x_segment_labels, self._knn, True, True,
        lambda t: graph_common.exp_inner_product_kernel(t, 5))
    return dmon_pool_loss(x.transpose(1, 2), adj, logits.transpose(1, 2), ~x_padding_mask)
