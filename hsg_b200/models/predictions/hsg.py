"""Loss head of HSG (hsg/models/predictions/hsg.py, `Hsg.losses` :78-227) with the three
pixel-to-prototype NCE terms evaluated in ONE pass over E x P^T.

The reference calls SegSortLoss three times on the same (embeddings, prototypes) with three
label sets -- image-level positives (:88-110), fine grouping (:120-136), coarse grouping
(:138-158) -- i.e. three [N,P] similarity matrices.  `nce_losses` hands the three label sets to
the fused kernel (`segsort_loss_multi`, K4) together; `losses` is the drop-in method with the
reference's return tuple: NCE terms, top-5 accuracy, the DMoN regulariser on the k-NN graph kernel
(:161-183) and the across-image centroid contrast (:185-225).
"""

import torch

from ...utils.general import common as common_utils
from ...utils.segsort import eval as segsort_eval
from ...utils.segsort import loss as segsort_loss

ORIGINALS = {}      # (module, class) -> the reference's `losses` (kept for unpatch / inspection)


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


def clustering_losses(self, datas, targets):
  """DMoN + collapse regularisation of both grouping levels and the centroid contrast across images,
  weighted and summed like the reference (:161-225); None when both are switched off."""
  total = None
  if self.dmon_loss is not None:                         # :161-183
    protos = datas['nd_prototype']
    mask = datas['nd_prototype_padding_mask']
    dmon_terms, collapse_terms = [], []
    for logits in (datas['coarsehrchy_nd_prototype_grouping_logit'], datas['finehrchy_nd_prototype_grouping_logit']):
      if 'nd_prototype_batch_index' in datas:
        d, c = self.dmon_loss(logits, protos, mask, datas['nd_prototype_batch_index'])
      else:                                              # the Cityscapes head passes no segment labels (hsg_cs.py:174-175)
        d, c = self.dmon_loss(logits, protos, mask)
      dmon_terms.append(d)
      collapse_terms.append(c)
    total = (sum(dmon_terms) + sum(collapse_terms)) * self.dmon_loss_weight
  if self.centroid_cont_loss is not None:                # :185-225
    img = targets['image_index'][datas['cluster_batch_index']]
    lo, hi = img.min().detach(), (img.max() + 1).detach()
    terms = []
    for prefix in ('coarse', 'fine'):
      tgt = targets[prefix + 'hrchy_nd_prototype_grouping_centroid']                 # [B', C, Q] of the whole batch
      b_all, _, q = tgt.shape
      tgt_rows = common_utils.normalize_embedding(tgt.permute(0, 2, 1).contiguous().flatten(0, 1))
      tgt_labels = torch.arange(tgt_rows.shape[0], dtype=torch.long, device=tgt_rows.device)
      mine = datas[prefix + 'hrchy_nd_prototype_grouping_centroid']
      mine_rows = common_utils.normalize_embedding(mine.permute(0, 2, 1).contiguous().flatten(0, 1))
      mine_labels = tgt_labels.view(b_all, q)[lo:hi].reshape(-1)                     # this GPU's images inside the batch
      terms.append(self.centroid_cont_loss(mine_rows, mine_labels, mine_labels, tgt_rows, tgt_labels))
    cont = sum(terms) * self.centroid_cont_loss_weight
    total = cont if total is None else total + cont
  return total


def _losses(self, datas, targets, split_graph_by_view):
  img_sim_loss, fine, coarse = nce_losses(self, datas, targets)
  hrchy_group_loss = fine
  if coarse is not None:
    hrchy_group_loss = coarse if hrchy_group_loss is None else hrchy_group_loss + coarse
  img_sim_acc = None
  if self.img_sim_loss is not None:                     # :104-118
    p_img = targets['image_index'][targets['prototype_batch_index']]
    p_lab = targets['prototype_instance_label'] * self.label_divisor + p_img
    img_sim_acc, _ = segsort_eval.top_k_ranking(targets['prototype'], p_lab, targets['prototype'], p_lab, 5)
  if not split_graph_by_view:
    datas = {k: v for k, v in datas.items() if k != 'nd_prototype_batch_index'}
  return img_sim_loss, hrchy_group_loss, clustering_losses(self, datas, targets), img_sim_acc


def losses(self, datas, targets={}):
  """Drop-in for `Hsg.losses` (predictions/hsg.py:78-227):
  (img_sim_loss, hrchy_group_loss, clustering_loss, img_sim_acc)."""
  return _losses(self, datas, targets, True)


def losses_cs(self, datas, targets={}):
  """The Cityscapes head (predictions/hsg_cs.py): identical except that the DMoN graph is not split by view
  (:173-175 pass no segment labels)."""
  return _losses(self, datas, targets, False)
