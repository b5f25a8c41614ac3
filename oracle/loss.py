"""Oracle: pixel-to-prototype NCE ("SegSort+") loss and the closed-form
backward passes (numpy restatement, test infrastructure).

Restates hsg/utils/segsort/loss.py:15-82 (``_calculate_log_likelihood``) and
:133-190 (``SegSortLoss``) of the reference.  The backward formulas are the
ones derived in SURVEY.md appendix A.1; ``gen_golden.py`` pins them against the
reference's own autograd.
"""

import numpy as np

from .ops import EPS


def nce_terms(embeddings, semantic_labels, instance_labels, prototypes,
              prototype_semantic_labels, concentration, group_mode='segsort+',
              dtype=np.float32, chunk=8192):
  """Per-pixel (numerator, denominator, own, uses_other_positives).

  S = exp(c * E P^T); own = S[i, inst_i]; pos = sum_{psem_j == sem_i} S_ij - own;
  num = pos if pos > 0 else own ('segsort+'), num = own otherwise;
  den = sum_{psem_j != sem_i} S_ij + num.        (loss.py:49-80)"""
  e = np.asarray(embeddings, dtype).reshape(-1, np.shape(embeddings)[-1])
  p = np.asarray(prototypes, dtype).reshape(-1, np.shape(prototypes)[-1])
  sem = np.asarray(semantic_labels).reshape(-1)
  inst = np.asarray(instance_labels).reshape(-1)
  psem = np.asarray(prototype_semantic_labels).reshape(-1)
  n = e.shape[0]
  num = np.empty(n, dtype)
  den = np.empty(n, dtype)
  own = np.empty(n, dtype)
  use = np.empty(n, bool)
  c = np.asarray(concentration, dtype)
  for s in range(0, n, chunk):
    sl = slice(s, min(n, s + chunk))
    sim = np.exp((e[sl] @ p.T) * c)                                   # :49-51
    rows = np.arange(sim.shape[0])
    o = sim[rows, inst[sl]]                                           # :58-59
    same = (sem[sl].reshape(-1, 1) == psem.reshape(1, -1))            # :62
    if group_mode == 'segsort+':
      pos = np.sum(sim * same.astype(dtype), axis=1, dtype=dtype) - o  # :64-66
      u = pos > 0
      nm = np.where(u, pos, o)                                        # :67-70
    else:
      u = np.zeros_like(o, bool)
      nm = o
    neg = np.sum(sim * (~same).astype(dtype), axis=1, dtype=dtype)    # :74-77
    num[sl], den[sl], own[sl], use[sl] = nm, neg + nm, o, u
  return num, den, own, use


def calculate_log_likelihood(embeddings, semantic_labels, instance_labels,
                             prototypes, prototype_semantic_labels,
                             concentration, group_mode='segsort+', dtype=np.float32):
  """[N,1] negative log-likelihood  -log(num/den)   (loss.py:80-82)."""
  num, den, _, _ = nce_terms(embeddings, semantic_labels, instance_labels,
                             prototypes, prototype_semantic_labels,
                             concentration, group_mode, dtype)
  return (-np.log(num / den)).reshape(-1, 1).astype(dtype)


def nce_condition(embeddings, semantic_labels, instance_labels, prototypes,
                  prototype_semantic_labels, concentration):
  """float64 condition number of the reference's numerator,
  kappa_i = (sum_{same} S_ij) / num_i  (>= 1).  The reference forms
  pos = (sum_{same} S) - own in float32 (loss.py:64-66), which cancels when the
  other positives are tiny next to the pixel's own prototype: its per-pixel
  loss then carries a relative error of about kappa_i * 2^-23 (seen: 5e-4 on 4
  of 2048 pixels of KAT3 between two float32 summation orders).  Tests allow
  |dl_i| <= rtol*|l_i| + kappa_i * 1e-6."""
  num, _, own, use = nce_terms(embeddings, semantic_labels, instance_labels, prototypes,
                               prototype_semantic_labels, concentration, 'segsort+', np.float64)
  return np.where(use, (num + own) / num, 1.0)


def segsort_loss(embeddings, semantic_labels, instance_labels, prototypes,
                 prototype_semantic_labels, concentration=10,
                 group_mode='segsort+', reduction='mean', dtype=np.float32):
  """SegSortLoss.forward (loss.py:149-190)."""
  ll = calculate_log_likelihood(embeddings, semantic_labels, instance_labels,
                                prototypes, prototype_semantic_labels,
                                concentration, group_mode, dtype)
  if reduction == 'mean':
    return ll.mean(dtype=dtype)
  if reduction == 'sum':
    return ll.sum(dtype=dtype)
  return ll


def segsort_loss_backward(embeddings, semantic_labels, instance_labels,
                          prototypes, prototype_semantic_labels, concentration,
                          grad_per_pixel, group_mode='segsort+', dtype=np.float64,
                          chunk=4096):
  """Closed-form (dE, dP) for  L = sum_i grad_per_pixel_i * l_i.

  dl_i/dS_ij = [j in Neg_i]/den + w_ij (1/den - 1/num),
  w_ij = [psem_j == sem_i, j != inst_i] if other positives exist else [j == inst_i];
  G = c * g_i * dl/dS o S;  dE = G P;  dP = G^T E.      """
  e = np.asarray(embeddings, dtype).reshape(-1, np.shape(embeddings)[-1])
  p = np.asarray(prototypes, dtype).reshape(-1, np.shape(prototypes)[-1])
  sem = np.asarray(semantic_labels).reshape(-1)
  inst = np.asarray(instance_labels).reshape(-1)
  psem = np.asarray(prototype_semantic_labels).reshape(-1)
  g = np.asarray(grad_per_pixel, dtype).reshape(-1)
  c = dtype(concentration)
  de = np.zeros_like(e)
  dp = np.zeros_like(p)
  for s in range(0, e.shape[0], chunk):
    sl = slice(s, min(e.shape[0], s + chunk))
    sim = np.exp((e[sl] @ p.T) * c)
    rows = np.arange(sim.shape[0])
    same = (sem[sl].reshape(-1, 1) == psem.reshape(1, -1))
    own_mask = np.zeros_like(same)
    own_mask[rows, inst[sl]] = True
    o = sim[rows, inst[sl]]
    if group_mode == 'segsort+':
      pos = (sim * same).sum(1) - o
      u = pos > 0
    else:
      pos = np.zeros_like(o)
      u = np.zeros_like(o, bool)
    nm = np.where(u, pos, o)
    den = (sim * ~same).sum(1) + nm
    w = np.where(u.reshape(-1, 1), same & ~own_mask, own_mask)
    dl = (~same) / den.reshape(-1, 1) + w * (1.0 / den - 1.0 / nm).reshape(-1, 1)
    gm = c * g[sl].reshape(-1, 1) * dl * sim
    de[sl] = gm @ p
    dp += gm.T @ e[sl]
  return de, dp


# --------------------------------------------------------------------------
# backward of prototype pooling / segment mean / normalize   
# --------------------------------------------------------------------------
def prototypes_backward(embeddings, labels, num_prototypes, grad_prototypes,
                        dtype=np.float64, eps=EPS):
  """d/dx of normalize(scatter_sum(x by label)) (segsort/common.py:30-39 +
  general/common.py:116-120): dL/ds_k = (g_k - p_k <p_k,g_k>)/||s_k|| when
  ||s_k|| >= eps else g_k/eps; dL/dx_i = dL/ds_{label_i}."""
  x = np.asarray(embeddings, dtype).reshape(-1, np.shape(embeddings)[-1])
  lab = np.asarray(labels).reshape(-1)
  g = np.asarray(grad_prototypes, dtype)
  s = np.zeros((num_prototypes, x.shape[1]), dtype)
  np.add.at(s, lab, x)
  nrm = np.sqrt((s * s).sum(1, keepdims=True))
  big = nrm >= eps
  p = s / np.where(big, nrm, eps)
  ds = np.where(big, (g - p * (p * g).sum(1, keepdims=True)) / np.where(big, nrm, 1.0), g / eps)
  return ds[lab]


def segment_mean_backward(index, grad_mean, dtype=np.float64):
  """d/dx of segment_mean (general/common.py:123-147): g_{idx_i}/max(cnt,1)."""
  idx = np.asarray(index).reshape(-1)
  g = np.asarray(grad_mean, dtype)
  cnt = np.bincount(idx, minlength=g.shape[0]).astype(dtype)
  cnt = np.where(cnt == 0, 1.0, cnt)
  return (g / cnt.reshape(-1, 1))[idx]


def normalize_backward(x, grad_out, dtype=np.float64, eps=EPS):
  """d/dx of normalize_embedding: (g - xhat <xhat,g>)/||x||  (g/eps below eps)."""
  x = np.asarray(x, dtype)
  g = np.asarray(grad_out, dtype)
  nrm = np.sqrt((x * x).sum(-1, keepdims=True))
  big = nrm >= eps
  xh = x / np.where(big, nrm, eps)
  return np.where(big, (g - xh * (xh * g).sum(-1, keepdims=True)) / np.where(big, nrm, 1.0), g / eps)


# --------------------------------------------------------------------------
# a15  Hsg.losses (NCE terms + accuracy)   hsg/models/predictions/hsg.py:78-160
# --------------------------------------------------------------------------
def top_k_accuracy(embeddings, labels, prototypes, prototype_labels, top_k):
  """hsg/utils/segsort/eval.py:9-52 (accuracy only)."""
  aff = np.asarray(embeddings, np.float32) @ np.asarray(prototypes, np.float32).T
  top = np.argsort(-aff, axis=1, kind='stable')[:, :top_k]
  hit = np.asarray(prototype_labels)[top] == np.asarray(labels).reshape(-1, 1)
  return np.float32(hit.astype(np.float32).mean())


def hsg_nce_losses(emb, cidx, batch_index, instance_label, image_index, protos, proto_batch, proto_inst,
                   fine_map, coarse_map, concentration, weights, label_divisor=2048):
  """The three pixel-to-prototype terms of Hsg.losses and the retrieval accuracy:
  (img_sim_loss, hrchy_group_loss, img_sim_acc).  predictions/hsg.py:88-160."""
  image_index = np.asarray(image_index, np.int64)
  inst = np.asarray(instance_label, np.int64) * label_divisor + image_index[batch_index]      # :91-96
  p_inst = np.asarray(proto_inst, np.int64) * label_divisor + image_index[proto_batch]       # :98-104
  img = segsort_loss(emb, inst, cidx, protos, p_inst, concentration) * weights[0]             # :105-111
  acc = top_k_accuracy(protos, p_inst, protos, p_inst, 5)                                     # :113-118
  fine = segsort_loss(emb, np.asarray(fine_map)[cidx], cidx, protos, fine_map, concentration) * weights[1]
  coarse = segsort_loss(emb, np.asarray(coarse_map)[cidx], cidx, protos, coarse_map, concentration) * weights[2]
  return img, fine + coarse, acc
