"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (the reference lives at /root/reference, which
does not exist on the GPU box):

    python oracle/gen_golden.py

Every fixture stores the inputs next to the reference's outputs, so nothing at
test time depends on torch's RNG.  Seeds follow the reference (235,
pyscripts/train/train.py:34-35).  Two CPU-only shims are applied from outside
(never by editing the reference): ``device.index`` is None on CPU
(hsg/utils/segsort/common.py:376) and ``scatter_gather.gather`` asserts CUDA
(hsg/models/utils.py:56,172-178).
"""

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('HSG_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

import hsg.utils.general.common as g_common          # noqa: E402
import hsg.utils.segsort.common as s_common          # noqa: E402
import hsg.utils.segsort.loss as s_loss              # noqa: E402
import hsg.models.utils as m_utils                   # noqa: E402


def _np(t):
  return t.detach().cpu().numpy()


def save(name, **arrays):
  os.makedirs(OUT, exist_ok=True)
  path = os.path.join(OUT, name + '.npz')
  np.savez_compressed(path, **arrays)
  print('%-28s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


_REFERENCE_SEGMENT_BY_KMEANS = s_common.segment_by_kmeans


def segment_by_kmeans_cpu(*args, **kwargs):
  """Reference segment_by_kmeans with the CPU shim `device.index or 0`
  (hsg/utils/segsort/common.py:376: `.device.index` is None on CPU and
  `N * None` raises).  The reference source is re-compiled in memory with that
  one expression patched; nothing is written anywhere."""
  import inspect
  text = inspect.getsource(_REFERENCE_SEGMENT_BY_KMEANS)
  patched = text.replace('cur_cluster_indices.device.index',
                         '(cur_cluster_indices.device.index or 0)')
  assert patched != text
  scope = dict(s_common.__dict__)
  exec(compile(patched, '<segment_by_kmeans+cpu-shim>', 'exec'), scope)
  return scope['segment_by_kmeans'](*args, **kwargs)


def main():
  torch.set_num_threads(8)

  # ---------------------------------------------------------------- a2
  torch.manual_seed(235)
  grids = [((16, 16), (448, 448)), ((6, 6), (14, 14)), ((4, 4), (28, 28)),
           ((6, 6), (512, 512)), ((12, 24), (1024, 2048)), ((2, 3), (4, 6)),
           ((3, 3), (12, 12)), ((5, 5), (7, 9)), ((1, 1), (14, 14)), ((3, 2), (10, 14))]
  arrays = {}
  for i, (k, hw) in enumerate(grids):
    lab = s_common.initialize_cluster_labels(list(k), hw, 'cpu')
    arrays['k%d' % i] = np.asarray(k)
    arrays['hw%d' % i] = np.asarray(hw)
    # store as row/col vectors (labels = y + (ymax+1)*x is rank-1 separable)
    arrays['lab_col0_%d' % i] = _np(lab[:, 0])
    arrays['lab_row0_%d' % i] = _np(lab[0, :])
    arrays['lab_sum_%d' % i] = np.asarray(int(lab.sum()))
  loc = s_common.generate_location_features((7, 9), 'cpu', 'float')
  arrays['loc_7_9'] = _np(loc)
  loc = s_common.generate_location_features((448, 448), 'cpu', 'float')
  arrays['loc_448_y'] = _np(loc[:, 0, 0])
  arrays['loc_448_x'] = _np(loc[0, :, 1])
  save('init_and_loc', n=np.asarray(len(grids)), **arrays)

  # ---------------------------------------------------------------- a1
  torch.manual_seed(235)
  x = torch.randn(64, 5, 34)
  x[3, 2] = 0.0                       # zero vector -> stays zero (eps branch)
  x[7, 1] = 1e-14                     # below eps
  save('normalize', x=_np(x), y=_np(g_common.normalize_embedding(x)))

  # ---------------------------------------------------------------- a4/a5/a6 KAT1
  torch.manual_seed(235)
  x = g_common.normalize_embedding(torch.randn(4096, 66))
  l0 = torch.randint(0, 16, (4096,))
  labs = [l0]
  protos = []
  lab = l0
  for _ in range(10):
    p = s_common.calculate_prototypes_from_labels(x, lab, 16)
    lab = s_common.find_nearest_prototypes(x, p)
    protos.append(p)
    labs.append(lab)
  final = s_common.kmeans_with_initial_labels(x, l0, 16, 10)
  assert torch.equal(final, lab)
  pf = s_common.calculate_prototypes_from_labels(x, final, 16)
  # torch CPU scatter_add_ visits rows in ascending order == np.add.at
  chk = np.zeros((16, 66), np.float32)
  np.add.at(chk, _np(final), _np(x))
  ref_sum = torch.zeros(16, 66).scatter_add_(0, final.view(-1, 1).expand(-1, 66), x)
  assert np.array_equal(chk, _np(ref_sum)), 'scatter order assumption broken'
  save('kmeans_flat_kat1', x=_np(x), labels=np.stack([_np(l) for l in labs]),
       prototypes=np.stack([_np(p) for p in protos]), final_prototypes=_np(pf))

  # well separated flat case: every implementation must agree exactly
  torch.manual_seed(235)
  centres = g_common.normalize_embedding(torch.randn(12, 40))
  assign = torch.randint(0, 12, (3000,))
  x = g_common.normalize_embedding(centres[assign] + 0.05 * torch.randn(3000, 40))
  l0 = torch.randint(0, 12, (3000,))
  labs = [l0]
  lab = l0
  for _ in range(8):
    p = s_common.calculate_prototypes_from_labels(x, lab, 12)
    lab = s_common.find_nearest_prototypes(x, p)
    labs.append(lab)
  save('kmeans_flat_separated', x=_np(x), labels=np.stack([_np(l) for l in labs]),
       prototypes=_np(s_common.calculate_prototypes_from_labels(x, lab, 12)))

  # prototypes with empty bins and explicit max_label, 4-D input
  torch.manual_seed(235)
  x = torch.randn(2, 6, 5, 18)
  lab = torch.randint(0, 9, (2, 6, 5))
  lab[lab == 4] = 5
  save('prototypes_empty_bins', x=_np(x), labels=_np(lab),
       p12=_np(s_common.calculate_prototypes_from_labels(x, lab, 12)),
       pauto=_np(s_common.calculate_prototypes_from_labels(x, lab)))

  # ---------------------------------------------------------------- a7 / a8
  torch.manual_seed(235)
  sem = torch.randint(0, 5, (500,))
  inst = torch.randint(0, 7, (500,)) * 3
  pl, ul = s_common.prepare_prototype_labels(sem, inst, 5)
  pl2, ul2 = s_common.prepare_prototype_labels(sem, inst)
  x = torch.randn(500, 10)
  idx = torch.randint(0, 9, (500,))
  idx[idx == 2] = 3
  save('labels_and_segment_mean', sem=_np(sem), inst=_np(inst), proto_labels=_np(pl),
       unique_inst=_np(ul), proto_labels_256=_np(pl2), unique_inst_256=_np(ul2),
       x=_np(x), idx=_np(idx), mean=_np(g_common.segment_mean(x, idx)))

  # ---------------------------------------------------------------- a3 KAT2
  torch.manual_seed(235)
  emb = torch.randn(2, 32, 12, 12)
  labels = torch.zeros(2, 12, 12, dtype=torch.long)
  for by in range(2):
    for bx in range(2):
      labels[:, by * 6:(by + 1) * 6, bx * 6:(bx + 1) * 6] = by * 2 + bx
  labels[0, :2, :] = 99
  res = segment_by_kmeans_cpu(emb, labels, [3, 3], ignore_index=99, iterations=5)
  save('segment_by_kmeans_kat2', emb=_np(emb), labels=_np(labels),
       out_emb=_np(res[0]), out_emb_loc=_np(res[1]), out_labels=_np(res[2]),
       out_cluster=_np(res[3]), out_batch=_np(res[4]))

  # no labels, model-style local features (location only), odd sizes, one iter more
  torch.manual_seed(235)
  emb = torch.randn(3, 16, 10, 14)
  loc = s_common.generate_location_features((10, 14), 'cpu', 'float') - 0.5
  loc = loc.unsqueeze(0).expand(3, 10, 14, 2)
  res = segment_by_kmeans_cpu(emb, None, [3, 2], local_features=loc, iterations=4)
  res0 = segment_by_kmeans_cpu(emb, None, [3, 2], iterations=4)
  for a, b in zip(res, res0):
    assert torch.equal(a, b)
  # 4 local-feature channels (location + colour-like), labels, ignore
  loc4 = torch.cat([loc, 0.3 * torch.randn(3, 10, 14, 2)], -1)
  labels = torch.randint(0, 3, (3, 10, 14)) * 2048 + torch.randint(0, 2, (3, 10, 14))
  labels[1, 4:, 5:] = 7000
  labels[2] = 7000                  # a whole image ignored
  res4 = segment_by_kmeans_cpu(emb, labels, [2, 3], local_features=loc4,
                               ignore_index=7000, iterations=3)
  save('segment_by_kmeans_misc', emb=_np(emb),
       a_emb=_np(res[0]), a_emb_loc=_np(res[1]), a_labels=_np(res[2]),
       a_cluster=_np(res[3]), a_batch=_np(res[4]),
       loc4=_np(loc4), labels4=_np(labels),
       b_emb=_np(res4[0]), b_emb_loc=_np(res4[1]), b_labels=_np(res4[2]),
       b_cluster=_np(res4[3]), b_batch=_np(res4[4]))

  # ---------------------------------------------------------------- a14 KAT3 (+ grads)
  torch.manual_seed(235)
  n, p_, d = 2048, 64, 32
  e = g_common.normalize_embedding(torch.randn(n, d))
  inst = torch.randint(0, p_, (n,))
  protos = s_common.calculate_prototypes_from_labels(e, inst, p_)
  psem = torch.randint(0, 20, (p_,))
  sem = psem[inst]
  loss = s_loss.SegSortLoss(16)(e, sem, inst, protos, psem)
  none = s_loss.SegSortLoss(16, reduction='none')(e, sem, inst, protos, psem)
  e_g = e.clone().requires_grad_(True)
  p_g = protos.clone().requires_grad_(True)
  s_loss.SegSortLoss(16)(e_g, sem, inst, p_g, psem).backward()
  plain = s_loss.SegSortLoss(16, group_mode='segsort', reduction='none')(e, sem, inst, protos, psem)
  save('nce_kat3', e=_np(e), inst=_np(inst), protos=_np(protos), psem=_np(psem), sem=_np(sem),
       loss=_np(loss), per_pixel=_np(none), de=_np(e_g.grad), dp=_np(p_g.grad),
       per_pixel_plain=_np(plain))

  # every prototype its own class -> every pixel takes the fallback branch;
  # prototypes NOT derived from the pixels; concentration 10
  torch.manual_seed(236)
  n, p_, d = 700, 37, 24
  e = g_common.normalize_embedding(torch.randn(n, d))
  protos = g_common.normalize_embedding(torch.randn(p_, d))
  inst = torch.randint(0, p_, (n,))
  psem = torch.arange(p_)
  psem[30:] = 5                      # a few shared classes
  sem = psem[inst]
  e_g = e.clone().requires_grad_(True)
  p_g = protos.clone().requires_grad_(True)
  w = torch.randn(n, 1)
  ll = s_loss.SegSortLoss(10, reduction='none')(e_g, sem, inst, p_g, psem)
  (ll * w).sum().backward()
  save('nce_fallback', e=_np(e), inst=_np(inst), protos=_np(protos), psem=_np(psem), sem=_np(sem),
       w=_np(w), per_pixel=_np(ll), de=_np(e_g.grad), dp=_np(p_g.grad))

  # ---------------------------------------------------------------- pooling backward
  torch.manual_seed(235)
  x = torch.randn(300, 12, requires_grad=True)
  lab = torch.randint(0, 11, (300,))
  lab[lab == 6] = 7
  g = torch.randn(14, 12)
  p = s_common.calculate_prototypes_from_labels(x, lab, 14)
  (p * g).sum().backward()
  dx_proto = x.grad.clone()
  x.grad = None
  gm = torch.randn(11, 12)
  m = g_common.segment_mean(x, lab)
  (m * gm).sum().backward()
  dx_mean = x.grad.clone()
  x.grad = None
  gn = torch.randn(300, 12)
  (g_common.normalize_embedding(x) * gn).sum().backward()
  save('pool_backward', x=_np(x), labels=_np(lab), g=_np(g), p=_np(p), dx_proto=_np(dx_proto),
       gm=_np(gm), mean=_np(m), dx_mean=_np(dx_mean), gn=_np(gn), dx_norm=_np(x.grad))

  # ---------------------------------------------------------------- a9 per-image(-pair) prototypes
  from hsg.models.embeddings.resnet_fcn_hsg import MultiviewResnetFcn, ResnetFcn
  torch.manual_seed(235)
  emb = torch.randn(4, 16, 8, 8)
  labels = torch.randint(0, 3, (4, 8, 8)) * 2048 + torch.randint(0, 2, (4, 8, 8))
  e_, el_, lab_, cidx_, bidx_ = segment_by_kmeans_cpu(emb, labels, [2, 2], iterations=3)
  pos = torch.randn(e_.shape[0], 16)
  fake_self = types.SimpleNamespace(label_divisor=2048, max_num_clusters=256)
  image_indices = torch.tensor([0, 0, 1, 1])
  mv = MultiviewResnetFcn._calculate_kmeans_prototypes(fake_self, e_, cidx_, bidx_, pos, lab_, image_indices)
  sv = ResnetFcn._calculate_kmeans_prototypes(fake_self, e_, cidx_, bidx_, pos, lab_)
  save('kmeans_prototypes', emb=_np(e_), cluster=_np(cidx_), batch=_np(bidx_), pos=_np(pos),
       labels=_np(lab_), image_indices=_np(image_indices),
       **{'mv%d' % i: _np(t) for i, t in enumerate(mv)},
       **{'sv%d' % i: _np(t) for i, t in enumerate(sv)})

  # ---------------------------------------------------------------- a12 hierarchy helpers
  torch.manual_seed(235)
  protos = torch.randn(3, 16, 20)
  glab = torch.randint(0, 5, (3, 20))
  pmask = torch.zeros(3, 20, dtype=torch.bool)
  pmask[0, 15:] = True
  pmask[2, 3:] = True
  coarse_n = ResnetFcn._collect_nd_coarser_prototype(fake_self, protos, glab, pmask, num_groups=6, normalized=True)
  coarse_m = ResnetFcn._collect_nd_coarser_prototype(fake_self, protos, glab, None, num_groups=None, normalized=False)
  pg = protos.clone().requires_grad_(True)
  gout = torch.randn(3, 16, 6)
  (ResnetFcn._collect_nd_coarser_prototype(fake_self, pg, glab, pmask, num_groups=6, normalized=True) * gout).sum().backward()
  pix_batch = torch.tensor([4, 4, 4, 7, 7, 9, 9, 9, 9, 4, 7])          # not sorted: the reference regroups
  pix_cidx = torch.randint(0, 20, (11,))
  pix_fine = ResnetFcn._collect_pixel_hierarchical_clustering_indices(fake_self, pix_cidx, pix_batch, glab)
  save('hierarchy', protos=_np(protos), glab=_np(glab), pmask=_np(pmask), coarse_norm=_np(coarse_n),
       coarse_mean=_np(coarse_m), gout=_np(gout), dprotos=_np(pg.grad), pix_batch=_np(pix_batch),
       pix_cidx=_np(pix_cidx), pix_fine=_np(pix_fine))

  # ---------------------------------------------------------------- a15 Hsg.losses (three NCE terms + accuracy)
  from hsg.models.predictions.hsg import Hsg
  ns = types.SimpleNamespace
  cfg = ns(train=ns(img_sim_loss_types='segsort', img_sim_concentration=16, img_sim_loss_weight=1.0,
                    fine_hrchy_loss_types='segsort', fine_hrchy_concentration=16, fine_hrchy_loss_weight=0.1,
                    coarse_hrchy_loss_types='segsort', coarse_hrchy_concentration=16, coarse_hrchy_loss_weight=0.1,
                    dmon_loss_types='dmon', dmon_knn=2, dmon_loss_weight=0.5,
                    centroid_cont_loss_types='segsort', centroid_cont_concentration=16,
                    centroid_cont_loss_weight=1.0),
           dataset=ns(semantic_ignore_index=255, num_classes=21), network=ns(label_divisor=2048))
  head = Hsg(cfg)
  torch.manual_seed(235)
  n_pix, n_proto, dim, n_img = 1500, 90, 32, 6              # 6 batch entries = 3 images x 2 views
  image_index = torch.tensor([0, 0, 1, 1, 2, 2])
  proto_batch = torch.sort(torch.randint(0, n_img, (n_proto,))).values
  proto_inst = torch.randint(0, 4, (n_proto,))
  cidx = torch.randint(0, n_proto, (n_pix,))
  emb = g_common.normalize_embedding(torch.randn(n_pix, dim)).requires_grad_(True)
  protos = g_common.normalize_embedding(torch.randn(n_proto, dim)).requires_grad_(True)
  fine_map = torch.randint(0, 24, (n_proto,))
  coarse_map = fine_map // 3
  cent_t = {k: torch.randn(3, dim, q) for k, q in (('fine', 8), ('coarse', 4))}
  cent_d = {k: torch.randn(3, dim, q).requires_grad_(True) for k, q in (('fine', 8), ('coarse', 4))}
  m_nodes = 20
  nd_proto = g_common.normalize_embedding(torch.randn(3, m_nodes, dim)).transpose(1, 2).contiguous()
  nd_mask = torch.zeros(3, m_nodes, dtype=torch.bool)
  nd_mask[0, 17:] = True
  nd_mask[2, 12:] = True
  nd_batch = torch.randint(0, 2, (3, m_nodes)) + 2 * torch.arange(3).view(3, 1)
  nd_fine = torch.softmax(torch.randn(3, 8, m_nodes), 1).requires_grad_(True)
  nd_coarse = torch.softmax(torch.randn(3, 4, m_nodes), 1).requires_grad_(True)
  datas = {'cluster_index': cidx, 'cluster_embedding': emb, 'cluster_batch_index': proto_batch[cidx],
           'cluster_instance_label': proto_inst[cidx],
           'finehrchy_nd_prototype_grouping_logit': nd_fine, 'coarsehrchy_nd_prototype_grouping_logit': nd_coarse,
           'nd_prototype': nd_proto, 'nd_prototype_batch_index': nd_batch, 'nd_prototype_padding_mask': nd_mask,
           'finehrchy_nd_prototype_grouping_centroid': cent_d['fine'],
           'coarsehrchy_nd_prototype_grouping_centroid': cent_d['coarse']}
  targets = {'image_index': image_index, 'prototype': protos, 'prototype_batch_index': proto_batch,
             'prototype_instance_label': proto_inst, 'finehrchy_mapping_index': fine_map,
             'coarsehrchy_mapping_index': coarse_map,
             'finehrchy_nd_prototype_grouping_centroid': cent_t['fine'],
             'coarsehrchy_nd_prototype_grouping_centroid': cent_t['coarse']}
  l_img, l_hr, l_cl, acc = head.losses(datas, targets)
  (l_img + l_hr + l_cl).backward()
  save('hsg_losses', image_index=_np(image_index), proto_batch=_np(proto_batch), proto_inst=_np(proto_inst),
       cidx=_np(cidx), emb=_np(emb), protos=_np(protos), fine_map=_np(fine_map), coarse_map=_np(coarse_map),
       cent_t_fine=_np(cent_t['fine']), cent_t_coarse=_np(cent_t['coarse']),
       cent_d_fine=_np(cent_d['fine']), cent_d_coarse=_np(cent_d['coarse']),
       img_sim_loss=_np(l_img), hrchy_group_loss=_np(l_hr), clustering_loss=_np(l_cl), accuracy=_np(acc),
       demb=_np(emb.grad), dprotos=_np(protos.grad), dcent_fine=_np(cent_d['fine'].grad),
       dcent_coarse=_np(cent_d['coarse'].grad), nd_proto=_np(nd_proto), nd_mask=_np(nd_mask), nd_batch=_np(nd_batch),
       nd_fine=_np(nd_fine), nd_coarse=_np(nd_coarse), dnd_fine=_np(nd_fine.grad), dnd_coarse=_np(nd_coarse.grad))

  # ---------------------------------------------------------------- 8f DMoN: k-NN affinity graph + loss
  import hsg.utils.graph.common as g_graph
  import hsg.utils.graph.loss as g_gloss
  torch.manual_seed(235)
  gx = g_common.normalize_embedding(torch.randn(3, 40, 24)).transpose(1, 2).contiguous()   # [B,C,n]: 40 nodes, 24 channels
  gpad = torch.zeros(3, 40, dtype=torch.bool)
  gpad[0, 33:] = True
  gpad[1, 1:] = True                      # a graph with a single valid node keeps its self loop
  gseg = torch.randint(0, 2, (3, 40)) * 7 + 3
  adj2 = g_graph.affinity_matrix_as_attention(gx, gpad, gseg, 2, True, True, lambda t_: g_graph.exp_inner_product_kernel(t_, 5))
  adj4 = g_graph.affinity_matrix_as_attention(gx, gpad, gseg, 4, True, False, lambda t_: g_graph.exp_inner_product_kernel(t_, 5))
  adj0 = g_graph.affinity_matrix_as_attention(gx, None, None, None, True, True)
  glog = torch.softmax(torch.randn(3, 6, 40), 1).requires_grad_(True)
  dl, cl = g_gloss.DMonLoss(adj_knn=2)(glog, gx, gpad, gseg)
  (dl + cl).backward()
  save('dmon', x=_np(gx), pad=_np(gpad), seg=_np(gseg), adj_knn2=_np(adj2), adj_knn4_values=_np(adj4), adj_noknn=_np(adj0),
       logits=_np(glog), dmon_loss=_np(dl), collapse_loss=_np(cl), dlogits=_np(glog.grad))

  # ---------------------------------------------------------------- a13 cross-GPU gather (2 "GPUs")
  m_utils.scatter_gather.gather = lambda xs, dev, dim=0: torch.cat(list(xs), dim)
  torch.manual_seed(235)
  ranks = []
  for r in range(2):
    emb = torch.randn(2, 16, 8, 8)
    labels = torch.randint(0, 3, (2, 8, 8)) * 2048 + torch.randint(0, 2, (2, 8, 8))
    e_, el_, lab_, cidx_, bidx_ = segment_by_kmeans_cpu(emb, labels, [2, 2], iterations=3)
    ranks.append((e_, el_, cidx_, bidx_ + 2 * r, lab_ // 2048, lab_ % 2048))
  outs = m_utils.gather_clustering_and_update_prototypes(
      [r[0] for r in ranks], [r[1] for r in ranks], [r[2] for r in ranks],
      [r[3] for r in ranks], [r[4] for r in ranks], [r[5] for r in ranks])
  fine = [torch.randint(0, 5, r[2].shape) for r in ranks]
  # mapping needs a function of ids: make level-2 ids a function of level-1 ids
  fine = [(o * 7 + 3) % 5 for o in outs[5]]
  mapping = m_utils.gather_and_update_cluster_mappings(list(outs[5]), fine)
  img = [torch.tensor([5, 5, 9, 2]), torch.tensor([9, 7, 7, 2])]
  reord = m_utils.gather_and_reorder_image_indices(img)
  arrays = {}
  for r in range(2):
    for j, nm in enumerate(['emb', 'emb_loc', 'cluster', 'batch', 'sem', 'inst']):
      arrays['r%d_%s' % (r, nm)] = _np(ranks[r][j])
    arrays['r%d_updated' % r] = _np(outs[5][r])
    arrays['r%d_fine' % r] = _np(fine[r])
    arrays['r%d_img' % r] = _np(img[r])
    arrays['r%d_img_reordered' % r] = _np(reord[r])
  save('gather_prototypes', prototypes=_np(outs[0][0]), prototypes_loc=_np(outs[1][0]),
       proto_sem=_np(outs[2][0]), proto_inst=_np(outs[3][0]), proto_batch=_np(outs[4][0]),
       mapping=_np(mapping[0]), **arrays)


def transformer_fixture():
  """a10/a11: TransformerClustering of the reference, small shapes, with its weights,
  outputs in eval mode and in train mode with dropout 0, and gradients (train mode)."""
  from hsg.models.embeddings.transformer_clusters import TransformerClustering
  torch.manual_seed(235)
  b, c, s, q, k = 3, 32, 24, 6, 4
  net = TransformerClustering(num_clusters=k, d_model=c, nhead=4, num_encoder_layers=2,
                              num_decoder_layers=2, dim_feedforward=2 * c, dropout=0.0)
  # make the BN layers non-trivial
  for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm1d):
      m.running_mean.normal_(0, 0.2)
      m.running_var.uniform_(0.5, 1.5)
      m.weight.data.uniform_(0.5, 1.5)
      m.bias.data.normal_(0, 0.2)
  src = torch.randn(b, c, s)
  pos = torch.randn(b, c, s)
  query = torch.randn(q, c)
  mask = torch.zeros(b, s, dtype=torch.bool)
  mask[0, 20:] = True
  mask[1, 9:] = True
  arrays = {'w__' + n.replace('.', '__'): _np(p).copy() for n, p in net.state_dict().items()}   # copy: BN buffers change below
  net.eval()
  with torch.no_grad():
    ev = net(src, mask, query, pos)
  net.train()
  src_g = src.clone().requires_grad_(True)
  tr = net(src_g, mask, query, pos)
  w = [torch.randn_like(t) for t in tr]
  sum((t * wi).sum() for t, wi in zip(tr, w)).backward()
  grads = {'g__' + n.replace('.', '__'): _np(p.grad) for n, p in net.named_parameters() if p.grad is not None}
  save('transformer_clustering', src=_np(src), pos=_np(pos), query=_np(query), mask=_np(mask),
       cfg=np.asarray([b, c, s, q, k]),
       **{'eval%d' % i: _np(t) for i, t in enumerate(ev)},
       **{'train%d' % i: _np(t) for i, t in enumerate(tr)},
       **{'w%d' % i: _np(t) for i, t in enumerate(w)},
       dsrc=_np(src_g.grad), **arrays, **grads)


def inference_fixture():
  """the prototype-bank writer and the nearest-neighbour label retrieval of the inference
  scripts, on small full-resolution-style inputs.  Runs the reference's own `generate_clusters`
  (hsg/models/embeddings/resnet_fcn.py:90-148, unbound, on a stub holding its four config fields),
  `calculate_prototypes_from_labels`, `find_majority_label_index` (pyscripts/inference/prototype.py:181-208),
  `load_memory_banks` (hsg/utils/segsort/others.py:11-41), `Segsort.predictions`
  (hsg/models/predictions/segsort.py:66-123) and `majority_label_from_topk` (segsort/eval.py:55-72)."""
  import tempfile
  import hsg.utils.segsort.eval as s_eval
  import hsg.utils.segsort.others as s_others
  from hsg.models.embeddings.resnet_fcn import ResnetFcn
  from hsg.models.predictions.segsort import Segsort
  torch.manual_seed(235)
  rng = np.random.RandomState(235)
  d, hp, wp, h, w = 16, 20, 24, 18, 21                 # padded image 20x24, real image 18x21 (pad right / bottom)
  n_class = 5
  centres = torch.nn.functional.normalize(torch.randn(n_class, d), dim=1)
  stub = types.SimpleNamespace(label_divisor=2048, semantic_ignore_index=255, kmeans_num_clusters=[4, 4],
                               kmeans_iterations=10)
  orig = s_common.segment_by_kmeans
  s_common.segment_by_kmeans = segment_by_kmeans_cpu     # CPU shim (device.index or 0), see above
  arrays = {}
  banks = []
  try:
    with tempfile.TemporaryDirectory() as tmp:
      for img in range(7):
        # ground truth: blocks of classes; embedding = class centre + noise, normalised per pixel
        gt = torch.from_numpy(rng.randint(0, n_class, size=(3, 3))).long()
        gt = gt.repeat_interleave(6, 0).repeat_interleave(7, 1)[:h, :w].contiguous()
        full = torch.zeros(hp, wp, dtype=torch.long)
        full[:h, :w] = gt
        emb = centres[full] + 2.0 * torch.randn(hp, wp, d) / d ** 0.5
        emb = g_common.normalize_embedding(emb).permute(2, 0, 1).unsqueeze(0).contiguous()
        fake = torch.full((1, hp, wp), 255, dtype=torch.long)
        fake[:, :h, :w] = 0
        out = ResnetFcn.generate_clusters(stub, emb, fake, fake.clone())
        protos = s_common.calculate_prototypes_from_labels(out['cluster_embedding'], out['cluster_index'])
        keep, proto_labels = s_common.find_majority_label_index(gt.unsqueeze(0), out['cluster_index'])
        arrays.update({'emb%d' % img: _np(emb), 'gt%d' % img: _np(gt), 'fake%d' % img: _np(fake),
                       'cluster_index%d' % img: _np(out['cluster_index']),
                       'cluster_embedding%d' % img: _np(out['cluster_embedding']),
                       'cluster_semantic_label%d' % img: _np(out['cluster_semantic_label']),
                       'cluster_instance_label%d' % img: _np(out['cluster_instance_label']),
                       'cluster_batch_index%d' % img: _np(out['cluster_batch_index']),
                       'protos%d' % img: _np(protos), 'proto_labels%d' % img: _np(proto_labels), 'keep%d' % img: _np(keep)})
        if img < 6:                                      # images 0-5 form the memory bank; image 6 is the query
          np.save(os.path.join(tmp, 'img%d.npy' % img), {'prototype': _np(protos), 'prototype_label': _np(proto_labels)})
        else:
          bank_p, bank_l = s_others.load_memory_banks(tmp)
          pred, topk = Segsort.predictions(None, out, {'semantic_memory_prototype': bank_p,
                                                       'semantic_memory_prototype_label': bank_l})
          arrays.update(bank_p=_np(bank_p), bank_l=_np(bank_l), pred=_np(pred), topk=_np(topk))
  finally:
    s_common.segment_by_kmeans = orig
  votes = torch.from_numpy(rng.randint(0, 7, size=(40, 20))).long()
  arrays.update(votes=_np(votes), votes_majority=_np(s_eval.majority_label_from_topk(votes)),
                votes_majority9=_np(s_eval.majority_label_from_topk(votes, 9)))
  save('inference_bank', cfg=np.asarray([d, hp, wp, h, w, 4, 4, 10, 2048, 255]), **arrays)


if __name__ == '__main__':
  if len(sys.argv) > 1 and sys.argv[1] == 'transformer':
    transformer_fixture()
  elif len(sys.argv) > 1 and sys.argv[1] == 'inference':
    inference_fixture()
  else:
    main()
    transformer_fixture()
    inference_fixture()
