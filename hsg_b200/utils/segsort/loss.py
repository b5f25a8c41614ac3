"""Drop-in operators for hsg/utils/segsort/loss.py (pixel-to-prototype NCE).

``SegSortLoss`` keeps the reference's constructor and call signature
(:133-190).  ``segsort_loss_multi`` evaluates several label sets on the same
(embeddings, prototypes) in one pass -- what Hsg.losses needs
(hsg/models/predictions/hsg.py:105,130,149) -- instead of three [N,P] passes.
"""

import torch
from torch.nn.modules.loss import _Loss

from ... import ops


def _calculate_log_likelihood(embeddings, semantic_labels, instance_labels, prototypes,
                              prototype_semantic_labels, concentration, group_mode):
  """[N,1] negative log-likelihood of each pixel (reference :15-82)."""
  ll = ops.nce_log_likelihood(embeddings, instance_labels, semantic_labels.reshape(1, -1), prototypes,
                              prototype_semantic_labels.reshape(1, -1), concentration, [group_mode])
  return ll.reshape(-1, 1)


def _reduce(ll, reduction):
  if reduction == 'mean':
    return torch.mean(ll)
  if reduction == 'sum':
    return torch.sum(ll)
  return ll


class SegSortLoss(_Loss):

  def __init__(self, concentration=10, group_mode='segsort+', size_average=None, reduce=None,
               reduction='mean'):
    super(SegSortLoss, self).__init__(size_average, reduce, reduction)
    self.concentration = concentration
    self.group_mode = group_mode

  def __repr__(self):
    return 'SegSortLoss(concentration={:.2f}, group_mode={})'.format(self.concentration, self.group_mode)

  def forward(self, embeddings, semantic_labels, instance_labels, prototypes,
              prototype_semantic_labels, prototype_weights=None):
    ll = _calculate_log_likelihood(embeddings, semantic_labels, instance_labels, prototypes,
                                   prototype_semantic_labels, self.concentration, self.group_mode)
    return _reduce(ll, self.reduction)


def segsort_loss_multi(embeddings, instance_labels, semantic_label_sets, prototypes,
                       prototype_semantic_label_sets, concentration, group_modes=None,
                       reduction='mean', num_prototypes=None):
  """One pass over E x P for several (semantic_labels, prototype_semantic_labels)
  pairs.  Returns a list with one loss per set (same values as calling
  SegSortLoss once per set).  num_prototypes (int64 device tensor [1], optional): only that many leading rows of
  `prototypes` exist -- the count stays on the device (pool_prototypes / exchange_prototypes_counted)."""
  sem = torch.stack([s.reshape(-1) for s in semantic_label_sets], 0)
  psem = torch.stack([s.reshape(-1) for s in prototype_semantic_label_sets], 0)
  if group_modes is None:
    group_modes = ['segsort+'] * sem.shape[0]
  ll = ops.nce_log_likelihood(embeddings, instance_labels, sem, prototypes, psem, concentration,
                              group_modes, num_prototypes=num_prototypes)
  return [_reduce(ll[s].reshape(-1, 1), reduction) for s in range(sem.shape[0])]
