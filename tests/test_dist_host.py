"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the prototype
exchange that replaces the reference's gather-everything-to-one-GPU step
(hsg/models/utils.py:127-217).  The local pooling stage is played by the
oracle here (CPU); the exchange / offset / ordering / backward logic under test
is the product code in hsg_b200/models/utils.py."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, golden_path, out_dir):
  import sys
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from hsg_b200.models import utils as mu
  from oracle import ops as o_ops, protos as o_protos
  g = dict(np.load(golden_path))
  pre = 'r%d_' % rank
  emb, emb_loc = g[pre + 'emb'], g[pre + 'emb_loc']
  cidx, bidx, sem, inst = g[pre + 'cluster'], g[pre + 'batch'], g[pre + 'sem'], g[pre + 'inst']
  # local stage (oracle on CPU): rank the rank's own (batch, cluster, sem, inst) tuples and pool
  res = o_protos.gather_clustering_and_update_prototypes([emb], [emb_loc], [cidx], [bidx], [sem], [inst])
  protos = torch.from_numpy(res[0]).requires_grad_(True)
  out = mu.exchange_prototypes(torch.from_numpy(res[5][0]), protos, torch.from_numpy(res[1]),
                               torch.from_numpy(res[2]), torch.from_numpy(res[3]),
                               torch.from_numpy(res[4]))
  # every rank's "loss" touches every prototype; d/d(local protos) must be the sum over ranks
  weight = torch.arange(out[0].numel(), dtype=torch.float32).view_as(out[0]) * (rank + 1)
  (out[0] * weight).sum().backward()
  # the packed single-collective form must give the same tensors and the same gradient
  protos2 = torch.from_numpy(res[0]).requires_grad_(True)
  out2 = mu.exchange_prototypes(torch.from_numpy(res[5][0]), protos2, torch.from_numpy(res[1]),
                                torch.from_numpy(res[2]), torch.from_numpy(res[3]),
                                torch.from_numpy(res[4]), capacity=64)
  (out2[0] * weight).sum().backward()
  for a_, b_ in zip(out, out2):
    assert torch.equal(a_.detach(), b_.detach())
  assert torch.equal(protos.grad, protos2.grad)
  np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), prototypes=out[0].detach().numpy(),
           prototypes_loc=out[1].numpy(), sem=out[2].numpy(), inst=out[3].numpy(), batch=out[4].numpy(),
           updated=out[5].numpy(), grad=protos.grad.numpy(), n_local=np.asarray(res[0].shape[0]))
  # flat k-means exchange: all-reduce of centroid sums gives the same sums on both ranks
  sums = torch.from_numpy(o_ops.scatter_sum(emb, cidx % 4, 4))
  dist.all_reduce(sums)
  np.save(os.path.join(out_dir, 'sums%d.npy' % rank), sums.numpy())
  dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_prototype_exchange_matches_reference_gather(tmp_path):
  golden_path = os.path.join(ROOT, 'tests', 'golden', 'gather_prototypes.npz')
  port = 29500 + (os.getpid() % 2000)
  mp.spawn(_worker, args=(2, port, golden_path, str(tmp_path)), nprocs=2, join=True)
  g = dict(np.load(golden_path))
  r = [dict(np.load(os.path.join(str(tmp_path), 'rank%d.npz' % k))) for k in range(2)]
  for k in range(2):
    np.testing.assert_allclose(r[k]['prototypes'], g['prototypes'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r[k]['prototypes_loc'], g['prototypes_loc'], rtol=1e-5, atol=1e-7)
    assert np.array_equal(r[k]['sem'], g['proto_sem'])
    assert np.array_equal(r[k]['inst'], g['proto_inst'])
    assert np.array_equal(r[k]['batch'], g['proto_batch'])
    assert np.array_equal(r[k]['updated'], g['r%d_updated' % k])
  # backward: sum over ranks of d(loss_r)/d(prototypes), sliced to the rank's own rows
  total = g['prototypes'].size
  w = np.arange(total, dtype=np.float32).reshape(g['prototypes'].shape) * 3.0     # (1 + 2)
  n0 = int(r[0]['n_local'])
  np.testing.assert_allclose(r[0]['grad'], w[:n0], rtol=1e-6)
  np.testing.assert_allclose(r[1]['grad'], w[n0:], rtol=1e-6)
  s0 = np.load(os.path.join(str(tmp_path), 'sums0.npy'))
  s1 = np.load(os.path.join(str(tmp_path), 'sums1.npy'))
  assert np.array_equal(s0, s1)


def test_sorted_triple_ranking_matches_reference_order():
  """The sort-based ranking that takes over when the presence table of the relabel kernel would not fit
  (hsg_b200/models/utils.py: _rank_triples_by_sort) gives the ids, labels and batch indices of the reference's
  two `unique` passes (hsg/models/utils.py:181-197, restated in oracle/protos.py)."""
  import sys
  sys.path.insert(0, ROOT)
  from hsg_b200.models import utils as mu
  from oracle import protos as o_protos
  rng = np.random.RandomState(235)
  n = 4000
  bidx = np.sort(rng.randint(3, 9, n)).astype(np.int64)             # images 3..8, rank-major
  cidx = (bidx - 3) * 7 + rng.randint(0, 7, n)                      # dense per-GPU cluster ids, some unused
  sem = rng.randint(0, 5, n).astype(np.int64)
  inst = rng.randint(0, 300, n).astype(np.int64)
  emb = rng.randn(n, 4).astype(np.float32)
  ref = o_protos.gather_clustering_and_update_prototypes([emb], [emb], [cidx], [bidx], [sem], [inst])
  packed = torch.from_numpy(sem) * mu._PACK + torch.from_numpy(inst)
  ids, pl, pb, cnt = mu._rank_triples_by_sort(torch.from_numpy(bidx), torch.from_numpy(cidx), packed)
  assert cnt == ref[2].shape[0]
  np.testing.assert_array_equal(ids.numpy(), ref[5][0])
  np.testing.assert_array_equal((pl // mu._PACK).numpy(), ref[2])
  np.testing.assert_array_equal((pl % mu._PACK).numpy(), ref[3])
  np.testing.assert_array_equal(pb.numpy(), ref[4])
