"""SyncBN over torch.distributed (hsg_b200/nn/sync_batchnorm.py): world_size-2 gloo run on CPU against
nn.BatchNorm1d on the concatenated batch -- forward, running statistics, gradients (the reference's own
SyncBN tests compare against nn.BatchNorm the same way, lib/nn/sync_batchnorm/tests/test_sync_batchnorm.py)."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
  import sys
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from hsg_b200.nn.sync_batchnorm import convert_model
  torch.manual_seed(7)
  net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.ReLU(), torch.nn.BatchNorm1d(5))
  net = convert_model(net)
  full = torch.randn(12, 6, generator=torch.Generator().manual_seed(3))
  w = torch.randn(12, 5, generator=torch.Generator().manual_seed(4))
  lo, hi = (0, 5) if rank == 0 else (5, 12)                       # uneven split
  x = full[lo:hi].clone().requires_grad_(True)
  y = net(x)
  (y * w[lo:hi]).sum().backward()
  np.savez(os.path.join(out_dir, 'r%d.npz' % rank), y=y.detach().numpy(), dx=x.grad.numpy(),
           rm=net[1].running_mean.numpy(), rv=net[1].running_var.numpy(), dw=net[3].weight.grad.numpy(),
           db=net[3].bias.grad.numpy(),
           dlin=net[0].weight.grad.numpy())
  dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sync_batchnorm_matches_batchnorm_on_the_whole_batch(tmp_path):
  port = 29700 + (os.getpid() % 2000)
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  torch.manual_seed(7)
  ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.ReLU(), torch.nn.BatchNorm1d(5))
  full = torch.randn(12, 6, generator=torch.Generator().manual_seed(3)).requires_grad_(True)
  w = torch.randn(12, 5, generator=torch.Generator().manual_seed(4))
  y = ref(full)
  (y * w).sum().backward()
  r = [dict(np.load(os.path.join(str(tmp_path), 'r%d.npz' % k))) for k in range(2)]
  np.testing.assert_allclose(np.concatenate([r[0]['y'], r[1]['y']]), y.detach().numpy(), rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(np.concatenate([r[0]['dx'], r[1]['dx']]), full.grad.numpy(), rtol=1e-4, atol=1e-6)
  for k in range(2):
    np.testing.assert_allclose(r[k]['rm'], ref[1].running_mean.numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r[k]['rv'], ref[1].running_var.numpy(), rtol=1e-5, atol=1e-7)
  # parameter gradients are per-rank partial sums (the data-parallel all-reduce adds them)
  np.testing.assert_allclose(r[0]['dw'] + r[1]['dw'], ref[3].weight.grad.numpy(), rtol=1e-4, atol=1e-6)
  np.testing.assert_allclose(r[0]['db'] + r[1]['db'], ref[3].bias.grad.numpy(), rtol=1e-4, atol=1e-6)
  # (the first layer's weights sit under two normalisations: their true gradient is ~0, compare absolutely)
  np.testing.assert_allclose(r[0]['dlin'] + r[1]['dlin'], ref[0].weight.grad.numpy(), rtol=1e-3, atol=1e-5)


def test_sync_batchnorm_single_process_equals_batchnorm():
  import sys
  sys.path.insert(0, ROOT)
  from hsg_b200.nn.sync_batchnorm import convert_model, SynchronizedBatchNorm
  torch.manual_seed(1)
  bn = torch.nn.BatchNorm1d(4)
  sbn = convert_model(torch.nn.Sequential(torch.nn.BatchNorm1d(4)))[0]
  assert isinstance(sbn, SynchronizedBatchNorm)
  x = torch.randn(3, 4, 7)
  np.testing.assert_allclose(sbn(x).detach().numpy(), bn(x).detach().numpy(), rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(sbn.running_var.numpy(), bn.running_var.numpy(), rtol=1e-5)
  sbn.eval(); bn.eval()
  np.testing.assert_allclose(sbn(x).detach().numpy(), bn(x).detach().numpy(), rtol=1e-5, atol=1e-6)
