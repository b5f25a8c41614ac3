"""cfg1 -- the drop-in claim on hardware: the reference's OWN training-step code
(`MultiviewResnetFcn.generate_clusters` hsg/models/embeddings/resnet_fcn_hsg.py:786-968 with
`_hierarchical_grouping` :580-681, `model_utils.gather_*` hsg/models/utils.py:41-240, `Hsg.forward` /
`Hsg.losses` hsg/models/predictions/hsg.py:78-265, driven as pyscripts/train/train.py:157-269 drives
them) runs UNPATCHED on CUDA and again over `hsg_b200.patch()`, from the same weights and the same
synthetic inputs, and every output is compared.

The reference comes from baseline/_ref (tools/install_reference.py; copied by __graft_entry__.build()).
Shape = BASELINE configs[0]: 4 images (2 samples x 2 views), D=128, 14x14 grid, k-means 6x6, T=10,
block oversegmentation with a few ignored pixels, hierarchy 256 -> 8 -> 4.

Bars.  Integers: identical.  Floats: 1e-5 of the max-norm, with two measured exceptions, both taken from
the reference itself in the same run (never asserted):
  * the reference is not bit-reproducible on a GPU (scatter_add_/index_add_ atomics) and, with the
    hierarchy losses on, its train-mode BatchNorm over 2x8 / 2x4 tokens amplifies that to 1e-2 in the
    coarse level from one run to the next: an output may differ from the reference by at most four times the
    largest difference between three runs of the reference;
  * run-to-run noise only exercises the reference's atomics, while any other fp32 implementation differs from it
    in the last bit everywhere: the yardstick also includes the reference's own response to a one-ulp change of
    its input floats (measured in the same run; up to three such runs next to five unperturbed ones).  Stage-2
    gradients (all five losses on) are compared as one vector through medians over those runs;
  * the NCE term forms `sum_same S - own` in fp32 (hsg/utils/segsort/loss.py:64-66), which cancels: the
    NCE losses are compared with the float64 value of the reference's own formula under the
    tolerance that formula admits in fp32 (1e-5 |l| + 1e-6 kappa per pixel); the
    synthetic input keeps kappa small (positives stay similar), so this is close to 1e-5 here.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402
import ref_step  # noqa: E402

pytestmark = pytest.mark.gpu

INT_KEYS = ['cluster_semantic_label', 'cluster_instance_label', 'cluster_index', 'cluster_batch_index',
            'finehrchy_cluster_index', 'coarsehrchy_cluster_index', 'nd_prototype_padding_mask',
            'nd_prototype_batch_index', 'nd_prototype_semantic_label', 'nd_prototype_instance_label',
            'cluster_index_by_image', 'finehrchy_nd_prototype_grouping_label',
            'coarsehrchy_nd_prototype_grouping_label']
FLOAT_KEYS = ['cluster_embedding', 'cluster_embedding_with_loc', 'nd_prototype',
              'finehrchy_nd_prototype_grouping_centroid', 'finehrchy_nd_prototype_grouping_logit',
              'finehrchy_nd_prototype_encoder_memory', 'coarsehrchy_nd_prototype_grouping_centroid',
              'coarsehrchy_nd_prototype_grouping_logit', 'coarsehrchy_nd_prototype_encoder_memory']
LABEL_KEYS_INT = ['image_index', 'prototype_semantic_label', 'prototype_instance_label', 'prototype_batch_index',
                  'finehrchy_mapping_index', 'coarsehrchy_mapping_index']
LABEL_KEYS_FLOAT = ['prototype', 'prototype_with_loc', 'finehrchy_prototype', 'coarsehrchy_prototype',
                    'finehrchy_prototype_with_loc', 'coarsehrchy_prototype_with_loc']
SPREAD_FACTOR = 4.0       # three runs give a small sample of the reference's spread; its tail is wider
LOSS_KEYS = ['img_sim_loss', 'hrchy_group_loss', 'clustering_loss', 'loss', 'accuracy']


def _rel(a, b, floor=0.0):
  a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
  return float(np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-30))


def _min_margin(outputs, level):
  """Smallest top-2 gap of the grouping soft assignment [B',Q,S] over the valid (unpadded) prototypes."""
  logit = outputs[level + '_nd_prototype_grouping_logit'].detach().double()
  top = torch.topk(logit, 2, dim=1).values
  gap = top[:, 0] - top[:, 1]
  return float(gap[~outputs['nd_prototype_padding_mask']].min())


def _nce_float64(e, sem, inst, p, psem, conc):
  """hsg/utils/segsort/loss.py:15-82 ('segsort+', mean) in float64, line by line, plus the fp32 tolerance the
  formula admits: per pixel 1e-5 |l| + 1e-6 kappa, kappa = (sum_same S) / num being the amplification of the
  `sum_same S - own` cancellation."""
  s = torch.exp(conc * e.detach().double() @ p.detach().double().t())
  same = sem.view(-1, 1) == psem.view(1, -1)
  own = s.gather(1, inst.view(-1, 1)).view(-1)
  cls = (s * same).sum(1)
  pos = cls - own
  num = torch.where(pos > 0, pos, own)
  neg = (s * (~same)).sum(1)
  loss = -torch.log(num / (neg + num))
  return float(loss.mean()), float((1e-5 * loss.abs() + 1e-6 * cls / num).mean())


def _nce_terms_float64(out, lab, cfg):
  """The NCE terms of Hsg.losses (hsg/models/predictions/hsg.py:86-157) on the reference's own embeddings,
  prototypes and label tables: {'img_sim_loss': (value, tolerance), 'hrchy_group_loss': (value, tolerance)}."""
  div, t = cfg.network.label_divisor, cfg.train
  e, p, cidx = out['cluster_embedding'], lab['prototype'], out['cluster_index']
  img = torch.gather(lab['image_index'], 0, out['cluster_batch_index'])
  pimg = torch.gather(lab['image_index'], 0, lab['prototype_batch_index'])
  v, tol = _nce_float64(e, out['cluster_instance_label'] * div + img, cidx, p,
                        lab['prototype_instance_label'] * div + pimg, float(t.img_sim_concentration))
  terms = {'img_sim_loss': (v * t.img_sim_loss_weight, tol * t.img_sim_loss_weight)}
  if t.fine_hrchy_loss_types != 'none':
    hv, ht = 0.0, 0.0
    for name, conc, w in (('finehrchy', t.fine_hrchy_concentration, t.fine_hrchy_loss_weight),
                          ('coarsehrchy', t.coarse_hrchy_concentration, t.coarse_hrchy_loss_weight)):
      psem = lab[name + '_mapping_index']
      v, tol = _nce_float64(e, torch.gather(psem, 0, cidx), cidx, p, psem, float(conc))
      hv, ht = hv + v * w, ht + tol * w
    terms['hrchy_group_loss'] = (hv, ht)
  return terms


def _runs(stage):
  import hsg_b200
  refenv.activate()
  dev = torch.device('cuda:0')
  torch.backends.cudnn.benchmark = False
  cfg = refenv.fresh_config(stage)
  hsg_b200.unpatch()
  ref_models = ref_step.build_models(cfg, dev)
  state = {k: v.detach().clone() for k, v in ref_models[0].state_dict().items()}
  # The two grouping levels end in an arg-max over (composed) soft assignments of a randomly initialised
  # transformer, and TransformerClustering orders its groups by a top-k over their largest logit
  # (hsg/models/embeddings/transformer_clusters.py:95-114): where two values tie to within float noise the
  # reference disagrees WITH ITSELF from run to run (group ids swap).  Use the first synthetic input (seed 235,
  # 236, ...) on which the reference's own decisions are certified: three runs of the reference give identical
  # integers, and every valid prototype's top-2 margin is above 1e-3 at both levels.
  for seed in range(235, 299):
    inputs = ref_step.make_inputs(dev, seed=seed)
    refs = [ref_step.run_step(ref_models[0], ref_models[1], inputs, dev) for _ in range(3)]
    margin = min(_min_margin(r[0], lvl) for r in refs for lvl in ('finehrchy', 'coarsehrchy'))
    stable = all(torch.equal(refs[0][0][k], r[0][k]) for r in refs[1:] for k in INT_KEYS) and \
        all(torch.equal(refs[0][1][k], r[1][k]) for r in refs[1:] for k in LABEL_KEYS_INT)
    print('stage %d seed %d: smallest grouping margin of the reference %.3e, three runs agree on every integer: %s'
          % (stage, seed, margin, stable))
    if margin > 1e-3 and stable:
      break
  else:
    pytest.fail('no synthetic input with certified grouping decisions')
  ref, ref2, ref3 = refs
  # Two more runs of the reference (five in all), and the reference's OWN response to a last-bit change of its inputs:
  # every input float times (1 +- 2^-23), i.e. moved by about one ulp.  Run-to-run noise only exercises the atomics;
  # any other fp32 implementation of the same operators (the reference's CPU path included) differs in the last bit
  # everywhere, and the train-mode BatchNorm over 2x8 / 2x4 tokens of the grouping levels amplifies either kind.
  perturbed = []
  refs = refs + [ref_step.run_step(ref_models[0], ref_models[1], inputs, dev) for _ in range(2)]
  for ps in range(8):
    g = torch.Generator().manual_seed(1000 + ps)
    moved = dict(inputs)
    for key in ('embedding', 'position_embedding'):
      sign = torch.randint(0, 2, inputs[key].shape, generator=g).to(dev).float() * 2 - 1
      moved[key] = inputs[key] * (1 + sign * 2.0 ** -23)
    run = ref_step.run_step(ref_models[0], ref_models[1], moved, dev)
    if all(torch.equal(ref[0][k], run[0][k]) for k in INT_KEYS) and \
        all(torch.equal(ref[1][k], run[1][k]) for k in LABEL_KEYS_INT):        # same decisions: a comparable run
      perturbed.append(run)
    if len(perturbed) == 3:
      break
  exact = _nce_terms_float64(ref[0], ref[1], cfg)
  launches = hsg_b200.load_library().hsg_launch_count()
  hsg_b200.patch()
  try:
    our_models = ref_step.build_models(cfg, dev, state=state)
    import hsg.models.embeddings.resnet_fcn_hsg as rf
    assert type(our_models[0].fine_hrchy_transformer).__module__.startswith('hsg_b200'), \
        'patch() must rebind the clustering transformer the reference model instantiates'
    assert rf.segsort_common.segment_by_kmeans.__module__.startswith('hsg_b200')
    ours = ref_step.run_step(our_models[0], our_models[1], inputs, dev)
    torch.cuda.synchronize()
  finally:
    hsg_b200.unpatch()
  launched = hsg_b200.load_library().hsg_launch_count() - launches
  return {'ref': ref, 'ref2': ref2, 'ref3': ref3, 'refs': refs, 'perturbed': perturbed, 'ours': ours, 'launched': launched, 'exact_nce': exact}


@pytest.fixture(scope='module')
def stage1():
  return _runs(1)


@pytest.fixture(scope='module')
def stage2():
  return _runs(2)


def _check_integers(r):
  for k in INT_KEYS:
    assert torch.equal(r['ref'][0][k], r['ours'][0][k]), k
  for k in LABEL_KEYS_INT:
    assert torch.equal(r['ref'][1][k], r['ours'][1][k]), k


def _check_floats(r, title):
  ref, ours = r['ref'], r['ours']
  # ref vs ref*: the largest difference between the first run of the reference and its four repeats and its (up to
  # three) runs on inputs moved by one ulp
  others = r['refs'][1:] + r['perturbed']
  report = []
  for src, keys in ((0, FLOAT_KEYS), (1, LABEL_KEYS_FLOAT)):
    for k in keys:
      report.append((k, _rel(ours[src][k], ref[src][k]), max(_rel(o[src][k], ref[src][k]) for o in others)))
  for k in LOSS_KEYS:
    if k in ref[2]:
      report.append((k, _rel(ours[2][k], ref[2][k]), max(_rel(o[2][k], ref[2][k]) for o in others)))
  print('\n%s\n%-50s %12s %12s' % (title, 'output', 'ours vs ref', 'ref vs ref*'))
  for k, a, b in report:
    print('%-50s %12.3e %12.3e' % (k, a, b))
  for k, (ex, tol) in r['exact_nce'].items():
    e_ours, e_ref = abs(float(ours[2][k]) - ex), abs(float(ref[2][k]) - ex)
    print('%s against float64 of the reference formula (%.9f): ours %.3e, reference fp32 %.3e, '
          'conditioning-aware tolerance %.3e' % (k, ex, e_ours, e_ref, tol))
    assert e_ours <= tol, (k, e_ours, tol)
  cancelling = ('img_sim_loss', 'hrchy_group_loss', 'loss')           # judged through float64 above
  bad = [(k, a, b) for k, a, b in report if k not in cancelling and a > max(1e-5, SPREAD_FACTOR * b)]
  assert not bad, 'outside max(1e-5, %g x reference run-to-run spread): %s' % (SPREAD_FACTOR, bad)
  return report


def _check_gradients(r, title, per_tensor=True):
  """Gradients of the total loss w.r.t. the input embeddings and every parameter that receives one.
  Differences are measured against max(|g|_max of that tensor, 1e-3 x the largest gradient of the step):
  the biases in front of a BatchNorm have a mathematically zero gradient, pure rounding noise."""
  ref, ref2, ref3, ours = r['ref'], r['ref2'], r['ref3'], r['ours']
  names = sorted(ref[3].keys())
  assert names == sorted(ours[3].keys())
  scale = max(float(ref[3][n_].abs().max()) for n_ in names)
  others = r['refs'][1:] + r['perturbed']
  rows = sorted(((_rel(ours[3][n_], ref[3][n_], 1e-3 * scale),
                  max(_rel(o[3][n_], ref[3][n_], 1e-3 * scale) for o in others), n_)
                 for n_ in names), reverse=True)
  print('\n%s: %d gradient tensors, largest |g| %.3e; largest differences (ours vs ref | ref vs ref):' % (title, len(names), scale))
  for a, b, n_ in rows[:8]:
    print('%-72s %10.3e %10.3e' % (n_, a, b))
  if per_tensor:
    bad = [(n_, a, b) for a, b, n_ in rows if a > max(1e-5, SPREAD_FACTOR * b)]
    assert not bad, bad[:5]
    return
  # all gradients as one vector: with every loss on, the reference's own gradients scatter by ~1e-1 per tensor from
  # run to run (train-mode BatchNorm over 2x8 / 2x4 tokens), far too noisy tensor by tensor
  def dist(x, y):
    num = sum(float((x[3][n_].double() - y[3][n_].double()).pow(2).sum()) for n_ in names)
    den = sum(float(y[3][n_].double().pow(2).sum()) for n_ in names)
    return (num / den) ** 0.5
  # Both distances are heavy-tailed (a near-tie flipping inside a BatchNorm batch of 8 tokens moves the whole vector), so
  # a single distance against the max of two is a coin toss.  Under the hypothesis "the patched step is one more run of
  # the reference" the distances ours<->ref_i are distributed like the pairwise distances ref_i<->ref_j: compare medians.
  refs = r['refs']
  d_ours = float(np.median([dist(ours, x) for x in refs]))
  pair = [dist(refs[i], refs[j]) for i in range(len(refs)) for j in range(len(refs)) if i != j]
  d_ref = float(np.median(pair))
  print('all gradients as one vector, relative l2 distance: median ours<->reference run %.3e over %d runs, median reference '
        'run<->run %.3e (min %.3e, max %.3e)' % (d_ours, len(refs), d_ref, min(pair), max(pair)))
  # the yardstick for "another fp32 implementation": the reference with every input float moved by one ulp
  moved = [dist(m, x) for m in r['perturbed'] for x in refs]
  d_ulp = float(np.median(moved)) if moved else 0.0
  print('reference with its inputs moved by one ulp (%d comparable runs): median distance to the unperturbed runs %.3e'
        % (len(r['perturbed']), d_ulp))
  # The distribution is bimodal, not just wide: a near-tie flipping inside a train-mode BatchNorm batch of 8 tokens puts
  # a run into one of two modes.  When the reference's five runs split 4-1 the median run<->run distance is the small
  # in-mode one; the patched step -- bit-reproducible, so always in the same mode -- is then either at that distance or at
  # the cross-mode one from the median run (one full-suite run in six landed there).  A deviation the reference shows
  # against ITSELF cannot count against the patched step, so the largest self-distance observed is admitted as it is
  # (no factor on it).
  d_self_max = max(pair + moved)
  print('largest distance of the reference to itself (repeats and one-ulp runs): %.3e' % d_self_max)
  assert d_ours <= max(1e-5, SPREAD_FACTOR * max(d_ref, d_ulp), d_self_max), (d_ours, d_ref, d_ulp, d_self_max)


def test_patched_step_runs_native_kernels(stage1):
  assert stage1['launched'] > 50, 'the patched step must run on libhsgb200.so kernels (%d launches)' % stage1['launched']


def test_stage1_integers_identical(stage1):
  _check_integers(stage1)


def test_stage1_floats(stage1):
  """image-similarity NCE only: the reference is deterministic up to atomics order here, so every output
  that the loss depends on sits at 1e-5 (the run-to-run column shows it)."""
  report = _check_floats(stage1, 'stage 1 (img-sim loss only)')
  strict = ['cluster_embedding', 'cluster_embedding_with_loc', 'nd_prototype', 'prototype', 'prototype_with_loc',
            'finehrchy_prototype', 'finehrchy_prototype_with_loc', 'finehrchy_nd_prototype_encoder_memory']
  bad = [(k, a) for k, a, b in report if k in strict and a > 1e-5]
  assert not bad, bad


def test_stage1_gradients(stage1):
  _check_gradients(stage1, 'stage 1')


def test_stage2_integers_identical(stage2):
  _check_integers(stage2)


def test_stage2_floats(stage2):
  _check_floats(stage2, 'stage 2 (all five loss terms)')


def test_stage2_gradients(stage2):
  _check_gradients(stage2, 'stage 2', per_tensor=False)

