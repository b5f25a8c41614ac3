"""float64 companions of two gradient fixtures: the UNMODIFIED reference run in double precision on the inputs
already stored in tests/golden/{nce_kat3,nce_fallback,hsg_losses}.npz.

    python oracle/gen_golden_f64.py          (build container only: needs /root/reference)

Why: a tolerance for "our fp32 gradient against the reference's fp32 gradient" has to be asserted; against the
float64 value of the reference's own formula it can be DERIVED -- ours must be as close to it as the reference's
own fp32 run is (tests/test_gpu_parity.py: test_nce_forward_backward, test_hsg_losses_*).  Writes
tests/golden/gradients_f64.npz; the existing fixtures are not touched.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('HSG_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

import hsg.utils.segsort.loss as s_loss              # noqa: E402


def load(name):
  with np.load(os.path.join(GOLD, name + '.npz')) as z:
    return {k: z[k] for k in z.files}


def t64(a):
  a = torch.from_numpy(np.asarray(a))
  return a.double() if a.dtype.is_floating_point else a


def main():
  out = {}
  for name, conc in (('nce_kat3', 16), ('nce_fallback', 10)):
    g = load(name)
    e = t64(g['e']).requires_grad_(True)
    p = t64(g['protos']).requires_grad_(True)
    pp = s_loss.SegSortLoss(conc, reduction='none')(e, t64(g['sem']), t64(g['inst']), p, t64(g['psem']))
    assert pp.dtype == torch.float64
    w = t64(g['w']) if 'w' in g else torch.full((g['e'].shape[0], 1), 1.0 / g['e'].shape[0], dtype=torch.float64)
    (pp * w).sum().backward()
    out[name + '__per_pixel'] = pp.detach().numpy()
    out[name + '__de'] = e.grad.numpy()
    out[name + '__dp'] = p.grad.numpy()
    print('%-14s fp32 reference vs its float64 run: per-pixel %.2e, dE %.2e, dP %.2e (max-norm relative)' % (
        name, np.abs(g['per_pixel'] - out[name + '__per_pixel']).max() / np.abs(out[name + '__per_pixel']).max(),
        np.abs(g['de'] - out[name + '__de']).max() / np.abs(out[name + '__de']).max(),
        np.abs(g['dp'] - out[name + '__dp']).max() / np.abs(out[name + '__dp']).max()))

  # Hsg.losses (hsg/models/predictions/hsg.py:78-227) on the hsg_losses fixture's inputs, in double
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
  g = load('hsg_losses')
  emb = t64(g['emb']).requires_grad_(True)
  protos = t64(g['protos']).requires_grad_(True)
  cent_f = t64(g['cent_d_fine']).requires_grad_(True)
  cent_c = t64(g['cent_d_coarse']).requires_grad_(True)
  nd_f = t64(g['nd_fine']).requires_grad_(True)
  nd_c = t64(g['nd_coarse']).requires_grad_(True)
  cidx, proto_batch, proto_inst = t64(g['cidx']), t64(g['proto_batch']), t64(g['proto_inst'])
  datas = {'cluster_index': cidx, 'cluster_embedding': emb, 'cluster_batch_index': proto_batch[cidx],
           'cluster_instance_label': proto_inst[cidx],
           'finehrchy_nd_prototype_grouping_logit': nd_f, 'coarsehrchy_nd_prototype_grouping_logit': nd_c,
           'nd_prototype': t64(g['nd_proto']), 'nd_prototype_batch_index': t64(g['nd_batch']),
           'nd_prototype_padding_mask': t64(g['nd_mask']),
           'finehrchy_nd_prototype_grouping_centroid': cent_f, 'coarsehrchy_nd_prototype_grouping_centroid': cent_c}
  targets = {'image_index': t64(g['image_index']), 'prototype': protos, 'prototype_batch_index': proto_batch,
             'prototype_instance_label': proto_inst, 'finehrchy_mapping_index': t64(g['fine_map']),
             'coarsehrchy_mapping_index': t64(g['coarse_map']),
             'finehrchy_nd_prototype_grouping_centroid': t64(g['cent_t_fine']),
             'coarsehrchy_nd_prototype_grouping_centroid': t64(g['cent_t_coarse'])}
  l_img, l_hr, l_cl, acc = head.losses(datas, targets)
  assert l_img.dtype == torch.float64, l_img.dtype
  (l_img + l_hr + l_cl).backward()
  vals = {'img_sim_loss': l_img, 'hrchy_group_loss': l_hr, 'clustering_loss': l_cl, 'demb': emb.grad, 'dprotos': protos.grad,
          'dcent_fine': cent_f.grad, 'dcent_coarse': cent_c.grad, 'dnd_fine': nd_f.grad, 'dnd_coarse': nd_c.grad}
  for k, v in vals.items():
    out['hsg_losses__' + k] = v.detach().numpy()
    ref32 = g[k].astype(np.float64)
    print('hsg_losses %-16s fp32 reference vs its float64 run: %.2e (max-norm relative)' % (
        k, np.abs(ref32 - out['hsg_losses__' + k]).max() / max(np.abs(out['hsg_losses__' + k]).max(), 1e-300)))
  np.savez_compressed(os.path.join(GOLD, 'gradients_f64.npz'), **out)
  print('wrote tests/golden/gradients_f64.npz')


if __name__ == '__main__':
  main()
