"""SyncBN kernels on the GPU (csrc/syncbn.cu behind hsg_b200/nn/sync_batchnorm.py):
  * one GPU: the statistics / apply kernels against nn.BatchNorm1d (values, running statistics, gradients,
    2-D and 3-D inputs, momentum=None, eval mode);
  * two GPUs (skipped on a one-GPU box): two NCCL ranks with an uneven split against nn.BatchNorm1d on the whole
    batch -- the one-process-per-GPU form of lib/nn/sync_batchnorm/batchnorm.py:55-118.  Also checks that
    hsg_b200.patch(sync_batchnorm=True) rebinds the reference's convert_model when baseline/_ref is present."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape,momentum', [((16, 256, 64), 0.1), ((12, 5), 0.1), ((3, 4, 7), None)])
def test_syncbn_kernels_equal_batchnorm(shape, momentum):
  sys.path.insert(0, ROOT)
  import hsg_b200
  from hsg_b200.nn.sync_batchnorm import convert_model, SynchronizedBatchNorm
  dev = torch.device('cuda:0')
  torch.manual_seed(1)
  c = shape[1]
  bn = torch.nn.BatchNorm1d(c, momentum=momentum).to(dev)
  with torch.no_grad():
    bn.weight.uniform_(0.5, 1.5)
    bn.bias.uniform_(-0.5, 0.5)
  sbn = convert_model(torch.nn.Sequential(torch.nn.BatchNorm1d(c, momentum=momentum))).to(dev)[0]
  sbn.load_state_dict(bn.state_dict())
  assert isinstance(sbn, SynchronizedBatchNorm)
  launches = hsg_b200.load_library().hsg_launch_count()
  for step in range(2):
    x = torch.randn(*shape, device=dev)
    w = torch.randn(*shape, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = bn(xa), sbn(xb)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    np.testing.assert_allclose(yb.detach().cpu().numpy(), ya.detach().cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(xb.grad.cpu().numpy(), xa.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(sbn.weight.grad.cpu().numpy(), bn.weight.grad.cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sbn.bias.grad.cpu().numpy(), bn.bias.grad.cpu().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sbn.running_mean.cpu().numpy(), bn.running_mean.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sbn.running_var.cpu().numpy(), bn.running_var.cpu().numpy(), rtol=1e-5, atol=1e-6)
    bn.zero_grad(); sbn.zero_grad()
  assert hsg_b200.load_library().hsg_launch_count() - launches == 8      # 2 steps x (stats, apply) x (fwd, bwd)
  bn.eval(); sbn.eval()
  x = torch.randn(*shape, device=dev)
  np.testing.assert_allclose(sbn(x).cpu().detach().numpy(), bn(x).cpu().detach().numpy(), rtol=1e-5, atol=1e-5)


def _worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.cuda.set_device(rank)
  dev = torch.device('cuda', rank)
  dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
  from hsg_b200.nn.sync_batchnorm import convert_model
  torch.manual_seed(7)
  net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.ReLU(), torch.nn.BatchNorm1d(5))
  net = convert_model(net).to(dev)
  full = torch.randn(12, 6, generator=torch.Generator().manual_seed(3))
  w = torch.randn(12, 5, generator=torch.Generator().manual_seed(4))
  lo, hi = (0, 5) if rank == 0 else (5, 12)                       # uneven split
  x = full[lo:hi].clone().to(dev).requires_grad_(True)
  y = net(x)
  (y * w[lo:hi].to(dev)).sum().backward()
  np.savez(os.path.join(out_dir, 'r%d.npz' % rank), y=y.detach().cpu().numpy(), dx=x.grad.cpu().numpy(),
           rm=net[1].running_mean.cpu().numpy(), rv=net[1].running_var.cpu().numpy(), dw=net[3].weight.grad.cpu().numpy(),
           db=net[3].bias.grad.cpu().numpy())
  dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_syncbn_two_nccl_ranks_match_batchnorm_on_the_whole_batch(tmp_path):
  if torch.cuda.device_count() < 2:
    pytest.skip('needs two GPUs (run with gpurun --gpus 2; result kept in profiles/r2_syncbn_nccl2.txt)')
  port = 29900 + (os.getpid() % 1000)
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  torch.manual_seed(7)
  ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.ReLU(), torch.nn.BatchNorm1d(5))
  full = torch.randn(12, 6, generator=torch.Generator().manual_seed(3)).requires_grad_(True)
  w = torch.randn(12, 5, generator=torch.Generator().manual_seed(4))
  y = ref(full)
  (y * w).sum().backward()
  r = [dict(np.load(os.path.join(str(tmp_path), 'r%d.npz' % k))) for k in range(2)]
  np.testing.assert_allclose(np.concatenate([r[0]['y'], r[1]['y']]), y.detach().numpy(), rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(np.concatenate([r[0]['dx'], r[1]['dx']]), full.grad.numpy(), rtol=1e-4, atol=1e-5)
  for k in range(2):
    np.testing.assert_allclose(r[k]['rm'], ref[1].running_mean.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r[k]['rv'], ref[1].running_var.numpy(), rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(r[0]['dw'] + r[1]['dw'], ref[3].weight.grad.numpy(), rtol=1e-4, atol=1e-5)
  np.testing.assert_allclose(r[0]['db'] + r[1]['db'], ref[3].bias.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_patch_can_rebind_the_reference_convert_model():
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import refenv
  if not refenv.available():
    pytest.skip('baseline/_ref is not installed')
  refenv.activate()
  import hsg_b200
  import lib.nn.sync_batchnorm.batchnorm as ref_bn
  original = ref_bn.convert_model
  hsg_b200.patch(sync_batchnorm=True)
  try:
    assert ref_bn.convert_model.__module__.startswith('hsg_b200')
    net = ref_bn.convert_model(torch.nn.Sequential(torch.nn.BatchNorm1d(4))).cuda()
    assert type(net[0]).__module__.startswith('hsg_b200')
    net(torch.randn(8, 4, device='cuda')).sum().backward()
  finally:
    hsg_b200.unpatch()
  assert ref_bn.convert_model is original
