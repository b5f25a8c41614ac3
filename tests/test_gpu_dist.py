"""Two NCCL ranks (skipped on a one-GPU box; run with `gpurun --gpus 2`, result kept in profiles/):
flat spherical k-means with its per-iteration all-reduce of exact int64 centroid sums must give labels
bit-identical to the one-GPU run, for an uneven split of the rows (reference
hsg/models/embeddings/clusters.py:30-42 -> hsg/utils/segsort/common.py:67-97), and the packed prototype
exchange must equal the multi-collective one over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _data():
  rng = np.random.RandomState(21)
  n, d, k = 300000, 258, 64
  x = rng.randn(n, d).astype(np.float32)
  x /= np.linalg.norm(x, axis=1, keepdims=True)
  return x, rng.randint(0, k, n).astype(np.int64), k


def _worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.cuda.set_device(rank)
  dev = torch.device('cuda', rank)
  dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
  from hsg_b200.models import utils as MU
  x, lab, k = _data()
  cut = 123457                                                     # uneven split
  lo, hi = (0, cut) if rank == 0 else (cut, x.shape[0])
  mine = MU.dist_kmeans_with_initial_labels(torch.from_numpy(x[lo:hi]).to(dev), torch.from_numpy(lab[lo:hi]).to(dev), k, 8)
  np.save(os.path.join(out_dir, 'labels%d.npy' % rank), mine.cpu().numpy())
  # packed exchange == multi-collective exchange
  g = torch.Generator().manual_seed(5 + rank)
  n_p = 7 + 5 * rank
  protos = torch.randn(n_p, 16, generator=g).to(dev).requires_grad_(True)
  ploc = torch.randn(n_p, 18, generator=g).to(dev)
  sem, inst, bat = [torch.randint(0, 9, (n_p,), generator=g).to(dev) for _ in range(3)]
  ids = torch.randint(0, n_p, (50,), generator=g).to(dev)
  a = MU.exchange_prototypes(ids, protos, ploc, sem, inst, bat)
  b = MU.exchange_prototypes(ids, protos, ploc, sem, inst, bat, capacity=16)
  ok = all(torch.equal(u.detach(), v.detach()) for u, v in zip(a, b))
  (b[0] * (rank + 1.0)).sum().backward()
  grad_ok = torch.allclose(protos.grad, torch.full_like(protos.grad, 3.0))
  # counted exchange (counts on the device, fixed capacity): the valid rows equal the packed exchange
  cap = 16
  pad = lambda t, fill: torch.cat([t.detach(), torch.full((cap - n_p,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=dev)], 0)
  protos_c = pad(protos, 7.0).requires_grad_(True)          # rows beyond the count hold junk on purpose
  count = torch.tensor([n_p], device=dev)
  c = MU.exchange_prototypes_counted(ids, protos_c, pad(ploc, 7.0), pad(sem, 3), pad(inst, 3), pad(bat, 3), count, cap)
  total = int(c[6])
  ok = ok and total == b[0].shape[0] and c[0].shape[0] == world * cap
  ok = ok and all(torch.equal(u.detach()[:total], v.detach()) for u, v in zip(c[:5], b[:5]))
  ok = ok and torch.equal(c[5], b[5]) and bool((c[0][total:] == 0).all()) and bool((c[2][total:] == -1).all())
  (c[0] * (rank + 1.0)).sum().backward()
  grad_ok = grad_ok and torch.allclose(protos_c.grad[:n_p], torch.full_like(protos_c.grad[:n_p], 3.0)) and \
      bool((protos_c.grad[n_p:] == 0).all())
  np.save(os.path.join(out_dir, 'ok%d.npy' % rank), np.asarray([ok, grad_ok]))
  dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_flat_kmeans_two_nccl_ranks_bit_identical_to_one_gpu(tmp_path):
  if torch.cuda.device_count() < 2:
    pytest.skip('needs two GPUs (run with gpurun --gpus 2; result kept in profiles/r2_nccl2_tests.txt)')
  port = 29300 + (os.getpid() % 1000)
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  sys.path.insert(0, ROOT)
  from hsg_b200.models import utils as MU
  x, lab, k = _data()
  dev = torch.device('cuda:0')
  one = MU.dist_kmeans_with_initial_labels(torch.from_numpy(x).to(dev), torch.from_numpy(lab).to(dev), k, 8, collective=False)
  two = np.concatenate([np.load(os.path.join(str(tmp_path), 'labels%d.npy' % r)) for r in range(2)])
  assert np.array_equal(one.cpu().numpy(), two), 'labels differ between 1 and 2 GPUs on %d rows' % int((one.cpu().numpy() != two).sum())
  for r in range(2):
    assert np.load(os.path.join(str(tmp_path), 'ok%d.npy' % r)).all()
