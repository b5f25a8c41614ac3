"""Loss head of HSG (hsg/models/predictions/hsg.py, `Hsg.losses` :78-227) with the three
pixel-to-prototype NCE terms evaluated in ONE pass over E x P^T.

The reference calls SegSortLoss three times on the same (embeddings, prototypes) with three
label sets -- image-level positives (:88-110), fine grouping (:120-136), coarse grouping
(:138-158) -- i.e. three [N,P] similarity matrices.  `nce_losses` hands the three label sets to
the fused kernel (`segsort_loss_multi`, K4) together; `losses` is the drop-in method: it keeps
the reference's return tuple and delegates the remaining regularisers (DMoN, centroid
contrast: SURVEY 8f "next") to the reference's own code.
"""

import torch

from ...utils.segsort import eval as segsort_eval
from ...utils.segsort import loss as segsort_loss

ORIGINALS = {}      # (module, class) -> the reference's `losses`, filled by hsg_b200.patch()


def nce_losses(self, datas, targets):
  """(img_sim_loss, fine_hrchy_loss, coarse_hrchy_loss), each already weighted, None when the
  term is switched off.  Terms that share concentration / group mode / reduction (every shipped
  recipe: bashscripts/coco/train.sh:51,123-126) go through the kernel together."""
  emb = datas['cluster_embedding']
  cidx = datas['cluster_index']
  protos = targets['prototype']
  terms = []                                            # (slot, module, weight, labels, prototype labels)
  if self.img_sim_loss is not None:
    div = self.label_divisor
    img = targets['image_index'][datas['cluster_batch_index']]
    p_img = targets['image_index'][targets['prototype_batch_index']]
    terms.append((0, self.img_sim_loss, self.img_sim_loss_weight,
                  datas['cluster_instance_label'] * div + img,
                  targets['prototype_instance_label'] * div + p_img))
  if self.fine_hrchy_loss is not None:
    plab = targets['finehrchy_mapping_index']
    terms.append((1, self.fine_hrchy_loss, self.fine_hrchy_loss_weight, plab[cidx], plab))
  if self.coarse_hrchy_loss is not None:
    plab = targets['coarsehrchy_mapping_index']
    terms.append((2, self.coarse_hrchy_loss, self.coarse_hrchy_loss_weight, plab[cidx], plab))

  out = [None, None, None]
  groups = {}
  for term in terms:
    mod = term[1]
    if isinstance(mod, segsort_loss.SegSortLoss):
      groups.setdefault((float(mod.concentration), mod.group_mode, mod.reduction), []).append(term)
    else:                                               # a loss object we do not know: call it as the reference does
      out[term[0]] = mod(emb, term[3], cidx, protos, term[4]) * term[2]
  for (conc, mode, reduction), members in groups.items():
    vals = segsort_loss.segsort_loss_multi(emb, cidx, [t[3] for t in members], protos, [t[4] for t in members],
                                           conc, [mode] * len(members), reduction)
    for t, v in zip(members, vals):
      out[t[0]] = v * t[2]
  return tuple(out)


class _WithoutNce(object):
  """`self` as the reference's `losses` sees it, with the three NCE terms switched off."""

  def __init__(self, module):
    object.__setattr__(self, '_m', module)

  def __getattr__(self, name):
    if name in ('img_sim_loss', 'fine_hrchy_loss', 'coarse_hrchy_loss'):
      return None
    return getattr(object.__getattribute__(self, '_m'), name)


def losses(self, datas, targets={}):
  """Drop-in for `Hsg.losses`: (img_sim_loss, hrchy_group_loss, clustering_loss, img_sim_acc)."""
  img_sim_loss, fine, coarse = nce_losses(self, datas, targets)
  hrchy_group_loss = fine
  if coarse is not None:
    hrchy_group_loss = coarse if hrchy_group_loss is None else hrchy_group_loss + coarse
  img_sim_acc = None
  if self.img_sim_loss is not None:                     # :104-118
    p_img = targets['image_index'][targets['prototype_batch_index']]
    p_lab = targets['prototype_instance_label'] * self.label_divisor + p_img
    img_sim_acc, _ = segsort_eval.top_k_ranking(targets['prototype'], p_lab, targets['prototype'], p_lab, 5)
  clustering_loss = None
  if self.dmon_loss is not None or self.centroid_cont_loss is not None:
    cls = type(self)
    original = None
    for c in cls.__mro__:
      original = ORIGINALS.get((c.__module__, c.__name__))
      if original is not None:
        break
    if original is None:
      raise RuntimeError('hsg_b200: the DMoN / centroid-contrast regularisers run through the reference\'s '
                         'Hsg.losses; call hsg_b200.patch() with the reference importable')
    clustering_loss = original(_WithoutNce(self), datas, targets)[2]
  return img_sim_loss, hrchy_group_loss, clustering_loss, img_sim_acc
