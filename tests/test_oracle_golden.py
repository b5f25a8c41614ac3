"""Pins the numpy oracle against outputs of the reference itself
(tests/golden/*.npz, made by oracle/gen_golden.py) and against the
known-answer values in SURVEY.md appendix A.4.  CPU only."""

import hashlib

import numpy as np

from oracle import ops, loss, protos

RTOL = 1e-5    # north_star: fp32 values within 1e-5 relative


def close(a, b, rtol=RTOL, atol=1e-6):
  np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def sha16(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_init_labels_and_location_features(golden):
  g = golden('init_and_loc')
  for i in range(int(g['n'])):
    k, hw = tuple(g['k%d' % i]), tuple(g['hw%d' % i])
    lab = ops.initialize_cluster_labels(k, hw)
    assert lab.dtype == np.int64 and lab.shape == hw
    assert np.array_equal(lab[:, 0], g['lab_col0_%d' % i]), (k, hw)
    assert np.array_equal(lab[0, :], g['lab_row0_%d' % i]), (k, hw)
    assert int(lab.sum()) == int(g['lab_sum_%d' % i])
  close(ops.generate_location_features((7, 9), 'float'), g['loc_7_9'], atol=1e-7)
  loc = ops.generate_location_features((448, 448), 'float')
  close(loc[:, 0, 0], g['loc_448_y'], atol=1e-7)
  close(loc[0, :, 1], g['loc_448_x'], atol=1e-7)
  # SURVEY A.4 known answer
  assert ops.initialize_cluster_labels((2, 3), (4, 6)).tolist() == [
      [0, 0, 2, 2, 4, 4], [0, 0, 2, 2, 4, 4], [1, 1, 3, 3, 5, 5], [1, 1, 3, 3, 5, 5]]
  assert ops.generate_location_features((3, 2), 'int').dtype == np.int64


def test_normalize(golden):
  g = golden('normalize')
  y = ops.normalize_embedding(g['x'])
  close(y, g['y'], atol=1e-7)
  assert np.all(y[3, 2] == 0)


def test_argmax_tie_rule():
  s = np.array([[1, 1, 0], [0, 2, 2], [0, 0, 0]], np.float32)
  assert ops.find_nearest_prototypes(np.eye(3, dtype=np.float32), s.T).tolist() == [0, 1, 0]


def test_kmeans_kat1_per_iteration(golden):
  g = golden('kmeans_flat_kat1')
  x, labels, protos_ref = g['x'], g['labels'], g['prototypes']
  for t in range(10):
    # teacher forced: reference labels in, one M-step + one E-step out
    p = ops.calculate_prototypes_from_labels(x, labels[t], 16)
    close(p, protos_ref[t], atol=1e-7)
    new = ops.find_nearest_prototypes(x, protos_ref[t])
    bad = np.nonzero(new != labels[t + 1])[0]
    if bad.size:                      # only near-ties may differ
      _, _, gap = ops.argmax_margins(x[bad], protos_ref[t])
      assert gap.max() < 2e-5
  # end to end this case is bit-identical to the reference
  final = ops.kmeans_with_initial_labels(x, labels[0], 16, 10)
  assert sha16(final) == '59513dc633417c12'
  assert np.bincount(final).tolist() == [245, 244, 238, 280, 252, 263, 261, 254,
                                         272, 259, 269, 237, 254, 257, 254, 257]
  pf = ops.calculate_prototypes_from_labels(x, final, 16)
  close(pf, g['final_prototypes'], atol=1e-7)
  assert abs(float(pf.sum()) - (-0.041798)) < 2e-5
  assert abs(float(pf[0, 0]) - (-0.0468399)) < 1e-6


def test_kmeans_separated(golden):
  g = golden('kmeans_flat_separated')
  final, trace = ops.kmeans_with_initial_labels(g['x'], g['labels'][0], 12, 8, return_trace=True)
  for t in range(8):
    assert np.array_equal(trace[t][2], g['labels'][t + 1])
  close(ops.calculate_prototypes_from_labels(g['x'], final, 12), g['prototypes'], atol=1e-7)
  # objective is non-decreasing across iterations
  obj = [ops.kmeans_objective(g['x'], tr[1], tr[2]) for tr in trace]
  assert all(b >= a - 1e-9 for a, b in zip(obj, obj[1:]))


def test_prototypes_empty_bins(golden):
  g = golden('prototypes_empty_bins')
  p12 = ops.calculate_prototypes_from_labels(g['x'], g['labels'], 12)
  close(p12, g['p12'], atol=1e-7)
  assert np.all(p12[4] == 0) and np.all(p12[9:] == 0)
  close(ops.calculate_prototypes_from_labels(g['x'], g['labels']), g['pauto'], atol=1e-7)


def test_labels_and_segment_mean(golden):
  g = golden('labels_and_segment_mean')
  pl, ul = ops.prepare_prototype_labels(g['sem'], g['inst'], 5)
  assert np.array_equal(pl, g['proto_labels']) and np.array_equal(ul, g['unique_inst'])
  pl, ul = ops.prepare_prototype_labels(g['sem'], g['inst'])
  assert np.array_equal(pl, g['proto_labels_256']) and np.array_equal(ul, g['unique_inst_256'])
  m = ops.segment_mean(g['x'], g['idx'])
  close(m, g['mean'], atol=1e-7)
  assert m.dtype == np.float32 and np.all(m[2] == 0)


def _check_segment(res, g, prefix):
  emb, emb_loc, lab, clu, bat = res
  close(emb, g[prefix + 'emb'], atol=1e-7)
  close(emb_loc, g[prefix + 'emb_loc'], atol=1e-6)
  assert np.array_equal(lab, g[prefix + 'labels'])
  assert np.array_equal(bat, g[prefix + 'batch'])
  assert np.array_equal(clu, g[prefix + 'cluster'])
  for a in (lab, clu, bat):
    assert a.dtype == np.int64


def test_segment_by_kmeans_kat2(golden):
  g = golden('segment_by_kmeans_kat2')
  res = ops.segment_by_kmeans(g['emb'], g['labels'], (3, 3), ignore_index=99, iterations=5)
  _check_segment(res, g, 'out_')
  assert all(r.shape[0] == 264 for r in res) and res[1].shape == (264, 34)
  assert abs(float(res[1].sum(dtype=np.float64)) - 17.662218) < 1e-3
  assert int(res[3].max()) == 41
  # SURVEY A.4 KAT2 hashes
  assert sha16(res[3]) == '5be8aa98f8cdfbfc'
  assert sha16(res[4]) == 'b5d269a12a932a86'
  assert sha16(res[2]) == '8ce55c738dc33b15'


def test_segment_by_kmeans_misc(golden):
  g = golden('segment_by_kmeans_misc')
  res = ops.segment_by_kmeans(g['emb'], None, (3, 2), iterations=4)
  _check_segment(res, g, 'a_')
  res = ops.segment_by_kmeans(g['emb'], g['labels4'], (2, 3), local_features=g['loc4'],
                              ignore_index=7000, iterations=3)
  _check_segment(res, g, 'b_')
  assert 2 not in res[4]          # the fully ignored image contributes nothing
  # gpu_id offset restates device.index (:376-377)
  res1 = ops.segment_by_kmeans(g['emb'], None, (3, 2), iterations=1, gpu_id=2)
  assert res1[4].min() == 6


def grad_close(de, dp, g, kappa):
  """Closed-form float64 grads vs the reference's float32 autograd: per pixel
  row within (1e-5 + 2e-6*kappa_i) of the row's largest entry (the reference's
  1/num term inherits the cancellation described in loss.nce_condition); dP,
  which mixes all pixels, within 1e-3 of its largest entry."""
  row_ref = np.abs(g['de']).max(1, keepdims=True)
  assert np.all(np.abs(de - g['de']) <= row_ref * (1e-5 + 2e-6 * kappa) + 1e-12)
  assert np.abs(dp - g['dp']).max() <= 1e-3 * np.abs(g['dp']).max()
  assert np.linalg.norm(dp - g['dp']) <= 1e-3 * np.linalg.norm(g['dp'])


def test_nce_kat3(golden):
  g = golden('nce_kat3')
  args = (g['e'], g['sem'], g['inst'], g['protos'], g['psem'])
  pp = loss.calculate_log_likelihood(*args, 16)
  # the reference's own float32 cancellation limits per-pixel agreement
  kappa = loss.nce_condition(*args, 16).reshape(-1, 1)
  assert np.all(np.abs(pp - g['per_pixel']) <= 1e-5 * np.abs(g['per_pixel']) + 1e-6 * kappa)
  assert np.mean(np.abs(pp - g['per_pixel']) <= 1e-5 * np.abs(g['per_pixel'])) > 0.99
  l = loss.segsort_loss(*args, concentration=16)
  assert abs(float(l) - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
  assert abs(float(l) - 4.7348456) < 1e-4
  assert abs(float(pp.sum(dtype=np.float64)) - 9696.96421) < 0.05
  close(loss.calculate_log_likelihood(*args, 16, group_mode='segsort'), g['per_pixel_plain'], rtol=2e-5)
  n = g['e'].shape[0]
  de, dp = loss.segsort_loss_backward(*args, 16, np.full(n, 1.0 / n))
  grad_close(de, dp, g, kappa)


def test_nce_fallback_branch(golden):
  g = golden('nce_fallback')
  args = (g['e'], g['sem'], g['inst'], g['protos'], g['psem'])
  _, _, _, use = loss.nce_terms(*args, 10)
  assert (~use).sum() > 0.5 * use.size          # fallback branch dominates here
  close(loss.calculate_log_likelihood(*args, 10), g['per_pixel'], rtol=2e-5)
  de, dp = loss.segsort_loss_backward(*args, 10, g['w'])
  grad_close(de, dp, g, loss.nce_condition(*args, 10).reshape(-1, 1))


def test_pool_backward(golden):
  g = golden('pool_backward')
  close(ops.calculate_prototypes_from_labels(g['x'], g['labels'], 14), g['p'], atol=1e-7)
  close(loss.prototypes_backward(g['x'], g['labels'], 14, g['g']), g['dx_proto'], rtol=1e-4, atol=1e-6)
  close(ops.segment_mean(g['x'], g['labels']), g['mean'], atol=1e-7)
  close(loss.segment_mean_backward(g['labels'], g['gm']), g['dx_mean'], rtol=1e-5, atol=1e-7)
  close(loss.normalize_backward(g['x'], g['gn']), g['dx_norm'], rtol=1e-4, atol=1e-6)


def test_kmeans_prototypes_per_image(golden):
  g = golden('kmeans_prototypes')
  for prefix, img in (('mv', g['image_indices']), ('sv', None)):
    out = protos.calculate_kmeans_prototypes(g['emb'], g['cluster'], g['batch'], g['pos'],
                                             g['labels'], img, 2048, 256)
    close(out[0], g[prefix + '0'], atol=1e-7)
    close(out[1], g[prefix + '1'], atol=1e-6)
    assert np.array_equal(out[2], g[prefix + '2'])
    assert np.array_equal(out[3], g[prefix + '3'])
    assert np.array_equal(out[4], g[prefix + '4'])
    assert np.array_equal(out[5], g[prefix + '5'])


def test_hierarchy_helpers(golden):
  g = golden('hierarchy')
  close(protos.collect_nd_coarser_prototype(g['protos'], g['glab'], g['pmask'], 6, True), g['coarse_norm'],
        rtol=1e-6, atol=1e-7)
  close(protos.collect_nd_coarser_prototype(g['protos'], g['glab'], None, None, False), g['coarse_mean'],
        rtol=1e-6, atol=1e-7)
  assert np.array_equal(protos.collect_pixel_hierarchical_clustering_indices(g['pix_cidx'], g['pix_batch'], g['glab']),
                        g['pix_fine'])


def test_hsg_losses_nce_terms(golden):
  g = golden('hsg_losses')
  img, hr, acc = loss.hsg_nce_losses(g['emb'], g['cidx'], g['proto_batch'][g['cidx']], g['proto_inst'][g['cidx']],
                                     g['image_index'], g['protos'], g['proto_batch'], g['proto_inst'],
                                     g['fine_map'], g['coarse_map'], 16, (1.0, 0.1, 0.1))
  assert abs(img - g['img_sim_loss']) <= 2e-5 * abs(g['img_sim_loss'])
  assert abs(hr - g['hrchy_group_loss']) <= 2e-5 * abs(g['hrchy_group_loss'])
  assert abs(acc - g['accuracy']) < 1e-6


def test_dmon_graph_and_loss(golden):
  from oracle import graph as o_graph
  g = golden('dmon')
  a = o_graph.exp_inner_product_kernel(g['x'], 5)
  assert np.array_equal(o_graph.knn_adjacency(a, g['pad'], g['seg'], 2, True, True), g['adj_knn2'])
  close(o_graph.knn_adjacency(a, g['pad'], g['seg'], 4, True, False), g['adj_knn4_values'], rtol=1e-6, atol=0)
  assert np.array_equal(o_graph.knn_adjacency(o_graph.exp_inner_product_kernel(g['x'], 5), None, None, None, True, True),
                        g['adj_noknn'])
  dl, cl = o_graph.dmon_pool_loss(g['adj_knn2'], np.transpose(g['logits'], (0, 2, 1)), ~g['pad'])
  assert abs(dl - g['dmon_loss']) < 1e-5 and abs(cl - g['collapse_loss']) < 1e-6


def test_cross_gpu_gather(golden):
  g = golden('gather_prototypes')
  ranks = [[g['r%d_%s' % (r, nm)] for nm in ('emb', 'emb_loc', 'cluster', 'batch', 'sem', 'inst')]
           for r in range(2)]
  out = protos.gather_clustering_and_update_prototypes(*[[rk[j] for rk in ranks] for j in range(6)])
  close(out[0], g['prototypes'], atol=1e-7)
  close(out[1], g['prototypes_loc'], atol=1e-7)
  assert np.array_equal(out[2], g['proto_sem'])
  assert np.array_equal(out[3], g['proto_inst'])
  assert np.array_equal(out[4], g['proto_batch'])
  for r in range(2):
    assert np.array_equal(out[5][r], g['r%d_updated' % r])
  table = protos.gather_and_update_cluster_mappings([g['r0_updated'], g['r1_updated']],
                                                    [g['r0_fine'], g['r1_fine']])
  assert np.array_equal(table, g['mapping'])
  re = protos.gather_and_reorder_image_indices([g['r0_img'], g['r1_img']])
  for r in range(2):                 # every rank receives the whole vector
    assert np.array_equal(re, g['r%d_img_reordered' % r])


def test_inference_prototype_bank_and_retrieval(golden, tmp_path):
  """generate_clusters -> prototype bank -> load_memory_banks -> nearest-neighbour labels."""
  from oracle import inference
  g = golden('inference_bank')
  d, hp, wp, h, w, ky, kx, iters, div, ignore = [int(v) for v in g['cfg']]
  for img in range(7):
    out = inference.generate_clusters(g['emb%d' % img], g['fake%d' % img], g['fake%d' % img], div, ignore, (ky, kx), iters)
    for key in ('cluster_index', 'cluster_semantic_label', 'cluster_instance_label', 'cluster_batch_index'):
      assert np.array_equal(out[key], g['%s%d' % (key, img)]), (key, img)
    close(out['cluster_embedding'], g['cluster_embedding%d' % img])
    protos, labels = inference.prototype_bank(out['cluster_embedding'], out['cluster_index'], g['gt%d' % img])
    close(protos, g['protos%d' % img])
    assert np.array_equal(labels, g['proto_labels%d' % img])
    keep, _ = inference.find_majority_label_index(g['gt%d' % img], out['cluster_index'])
    assert np.array_equal(keep, g['keep%d' % img])
    if img < 6:
      np.save(str(tmp_path / ('img%d.npy' % img)), {'prototype': protos, 'prototype_label': labels})
  bank_p, bank_l = inference.load_memory_banks(str(tmp_path))
  close(bank_p, g['bank_p'])
  assert np.array_equal(bank_l, g['bank_l'])
  pred, topk = inference.predictions(out['cluster_embedding'], out['cluster_index'], g['bank_p'], g['bank_l'])
  assert np.array_equal(pred, g['pred']) and np.array_equal(topk, g['topk'])
  assert np.array_equal(inference.majority_label_from_topk(g['votes']), g['votes_majority'])
  assert np.array_equal(inference.majority_label_from_topk(g['votes'], 9), g['votes_majority9'])
