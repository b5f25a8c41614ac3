"""Synchronised batch normalisation for one process per GPU.

The reference synchronises BN statistics across its DataParallel replicas with a master/slave
rendezvous over Python queues plus `ReduceAddCoalesced` / `Broadcast` of a [2,C] tensor per layer
(lib/nn/sync_batchnorm/batchnorm.py:55-118, comm.py:96-127) -- including the 28 BatchNorm1d layers
of the two clustering transformers (pyscripts/train/train.py:100-102).  With one process per GPU
the same statistics are ONE `all_reduce` of [2C+1] floats per layer over torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests), forward and backward, with one statistics kernel
before it and one apply kernel after it (csrc/syncbn.cu) instead of ~15 eager launches per layer.
`hsg_b200.patch(sync_batchnorm=True)` rebinds `lib.nn.sync_batchnorm.batchnorm.convert_model` (and makes
`patch_replication_callback` a no-op) for torchrun-style launches; the reference's thread-per-GPU DataParallel
keeps its own implementation by default, because replicas inside ONE process do not form a process group.

    model = convert_model(model)        # same name and behaviour as lib.nn.sync_batchnorm.convert_model

Numerics follow the reference / torch: biased variance for normalisation, unbiased for the running
estimate, momentum update of the running statistics, eval mode uses the running statistics.
"""

import torch
import torch.distributed as dist
from torch import nn


def _reduce(t, group):
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
  return t


def _kernels(x):
  """The two CUDA kernels either side of the all-reduce (csrc/syncbn.cu); CPU tensors -- the gloo host-logic
  tests -- take the eager expressions below."""
  return x.is_cuda


def _launch(fn_name, *args):
  from .. import _lib
  _lib.check(getattr(_lib.load(), fn_name)(*args), fn_name)


def _p(t):
  import ctypes
  return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
  import ctypes
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _SyncBatchNormFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, weight, bias, eps, group):
    # x: [N, C, L]; statistics over every dim but 1, over every rank
    b, c, l = x.shape
    xf = x.float().contiguous()
    stats = torch.empty((2 * c + 1,), dtype=torch.float32, device=x.device)
    stats[-1] = b * l
    if _kernels(xf):
      with torch.cuda.device(x.device):
        _launch('hsg_bn_stats_f32', _p(xf), None, None, None, b, c, l, _p(stats), _stream())
    else:
      stats[:c] = xf.sum((0, 2))
      stats[c:2 * c] = (xf * xf).sum((0, 2))
    _reduce(stats, group)
    count = stats[-1]
    mean = stats[:c] / count
    var = stats[c:2 * c] / count - mean * mean                     # biased
    inv = torch.rsqrt(var.clamp_min(0) + eps)
    w = weight.float().contiguous() if weight is not None else None
    bs = bias.float().contiguous() if bias is not None else None
    if _kernels(xf):
      out = torch.empty_like(xf)
      with torch.cuda.device(x.device):
        _launch('hsg_bn_apply_f32', _p(xf), None, _p(mean.contiguous()), _p(inv.contiguous()), _p(w), _p(bs), None, 0.0,
                b, c, l, _p(out), _stream())
    else:
      out = (xf - mean.view(1, c, 1)) * inv.view(1, c, 1)
      if w is not None:
        out = out * w.view(1, c, 1) + bs.view(1, c, 1)
    ctx.save_for_backward(xf, mean, inv, weight)
    ctx.group, ctx.count = group, count
    ctx.mark_non_differentiable(mean, var, count)
    return out.to(x.dtype), mean, var, count

  @staticmethod
  def backward(ctx, gout, _gm, _gv, _gc):
    xf, mean, inv, weight = ctx.saved_tensors
    b, c, l = xf.shape
    g = gout.float().contiguous()
    red = torch.empty((2 * c,), dtype=torch.float32, device=xf.device)   # sums of dy and of dy * xhat
    if _kernels(xf):
      with torch.cuda.device(xf.device):
        _launch('hsg_bn_stats_f32', _p(xf), _p(g), _p(mean.contiguous()), _p(inv.contiguous()), b, c, l, _p(red), _stream())
    else:
      xhat = (xf - mean.view(1, c, 1)) * inv.view(1, c, 1)
      red[:c] = g.sum((0, 2))
      red[c:] = (g * xhat).sum((0, 2))
    gb, gw = red[:c].clone(), red[c:].clone()                            # local parameter gradients
    _reduce(red, ctx.group)
    w = weight.float().contiguous() if weight is not None else None
    if _kernels(xf):
      gx = torch.empty_like(xf)
      with torch.cuda.device(xf.device):
        _launch('hsg_bn_apply_f32', _p(xf), _p(g), _p(mean.contiguous()), _p(inv.contiguous()), _p(w), None, _p(red),
                float(1.0 / float(ctx.count)), b, c, l, _p(gx), _stream())
    else:
      gamma = w if w is not None else torch.ones_like(inv)
      xhat = (xf - mean.view(1, c, 1)) * inv.view(1, c, 1)
      gx = (g - (red[:c] / ctx.count).view(1, c, 1) - xhat * (red[c:] / ctx.count).view(1, c, 1)) * (gamma * inv).view(1, c, 1)
    return gx.to(gout.dtype), (gw if weight is not None else None), (gb if weight is not None else None), None, None


class SynchronizedBatchNorm(nn.modules.batchnorm._BatchNorm):
  """BatchNorm{1,2,3}d whose batch statistics span every rank of `process_group`."""

  def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, process_group=None, track_running_stats=True):
    super().__init__(num_features, eps, momentum, affine, track_running_stats)
    self.process_group = process_group

  def _check_input_dim(self, x):
    if x.dim() < 2:
      raise ValueError('expected at least 2D input (got {}D input)'.format(x.dim()))

  def forward(self, x):
    self._check_input_dim(x)
    if not self.training and self.running_mean is not None:
      return nn.functional.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, False, 0.0,
                                      self.eps)
    shape = x.shape
    x3 = x.reshape(shape[0], shape[1], -1)
    out, mean, var, count = _SyncBatchNormFn.apply(x3, self.weight, self.bias, self.eps, self.process_group)
    if self.track_running_stats and self.running_mean is not None:
      with torch.no_grad():
        self.num_batches_tracked += 1
        # momentum=None: cumulative moving average, as nn.BatchNorm
        m = 1.0 / float(self.num_batches_tracked) if self.momentum is None else self.momentum
        unbiased = var * (count / (count - 1).clamp_min(1))
        self.running_mean.mul_(1 - m).add_(mean.to(self.running_mean.dtype), alpha=m)
        self.running_var.mul_(1 - m).add_(unbiased.to(self.running_var.dtype), alpha=m)
    return out.reshape(shape)


SynchronizedBatchNorm1d = SynchronizedBatchNorm2d = SynchronizedBatchNorm3d = SynchronizedBatchNorm


def convert_model(module, process_group=None):
  """Replace every BatchNorm{1,2,3}d in `module` (recursively) by a SynchronizedBatchNorm carrying the same
  parameters and running statistics; reference lib/nn/sync_batchnorm/batchnorm.py:353-393."""
  if isinstance(module, nn.DataParallel):            # reference :366-370: convert what the wrapper holds
    inner = convert_model(module.module, process_group)
    return nn.DataParallel(inner, device_ids=module.device_ids)
  mod = module
  if isinstance(module, nn.modules.batchnorm._BatchNorm) and not isinstance(module, SynchronizedBatchNorm):
    mod = SynchronizedBatchNorm(module.num_features, module.eps, module.momentum, module.affine, process_group,
                                module.track_running_stats)
    mod.running_mean = module.running_mean
    mod.running_var = module.running_var
    mod.num_batches_tracked = module.num_batches_tracked
    if module.affine:
      mod.weight.data = module.weight.data.clone().detach()
      mod.bias.data = module.bias.data.clone().detach()
    mod.train(module.training)
  for name, child in module.named_children():
    mod.add_module(name, convert_model(child, process_group))
  return mod


def patch_replication_callback(data_parallel):
  """lib/nn/sync_batchnorm/replicate.py:69-94 hooks the replicate() of a thread-per-GPU DataParallel so that the
  replicas find their master; with one process per GPU there are no replicas to connect: a no-op."""
  return data_parallel
