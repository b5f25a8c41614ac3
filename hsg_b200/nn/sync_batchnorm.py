"""Synchronised batch normalisation for one process per GPU.

The reference synchronises BN statistics across its DataParallel replicas with a master/slave
rendezvous over Python queues plus `ReduceAddCoalesced` / `Broadcast` of a [2,C] tensor per layer
(lib/nn/sync_batchnorm/batchnorm.py:55-118, comm.py:96-127) -- including the 28 BatchNorm1d layers
of the two clustering transformers (pyscripts/train/train.py:100-102).  With one process per GPU
the same statistics are ONE `all_reduce` of [2C+1] floats per layer over torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests), forward and backward.

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


class _SyncBatchNormFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, weight, bias, eps, group):
    # x: [N, C, *]; statistics over every dim but 1, over every rank
    c = x.shape[1]
    dims = [d for d in range(x.dim()) if d != 1]
    xf = x.float()
    stats = torch.cat([xf.sum(dims), (xf * xf).sum(dims), xf.new_tensor([xf.numel() / c])])
    _reduce(stats, group)
    count = stats[-1]
    mean = stats[:c] / count
    var = stats[c:2 * c] / count - mean * mean                     # biased
    inv = torch.rsqrt(var.clamp_min(0) + eps)
    shape = [1, c] + [1] * (x.dim() - 2)
    xhat = (xf - mean.view(shape)) * inv.view(shape)
    out = xhat
    if weight is not None:
      out = out * weight.float().view(shape) + bias.float().view(shape)
    ctx.save_for_backward(xhat, inv, weight)
    ctx.group, ctx.count, ctx.dims, ctx.shape = group, count, dims, shape
    ctx.mark_non_differentiable(mean, var, count)
    return out.to(x.dtype), mean, var, count

  @staticmethod
  def backward(ctx, gout, _gm, _gv, _gc):
    xhat, inv, weight = ctx.saved_tensors
    c = xhat.shape[1]
    g = gout.float()
    gw = (g * xhat).sum(ctx.dims)
    gb = g.sum(ctx.dims)
    red = _reduce(torch.cat([gb, gw]), ctx.group)                  # global sums of dy and dy * xhat
    gamma = weight.float() if weight is not None else torch.ones_like(inv)
    mean_g = (red[:c] / ctx.count).view(ctx.shape)
    mean_gx = (red[c:] / ctx.count).view(ctx.shape)
    gx = (g - mean_g - xhat * mean_gx) * (gamma * inv).view(ctx.shape)
    return gx.to(gout.dtype), (gw if weight is not None else None), (gb if weight is not None else None), None, None


class SynchronizedBatchNorm(nn.modules.batchnorm._BatchNorm):
  """BatchNorm{1,2,3}d whose batch statistics span every rank of `process_group`."""

  def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, process_group=None):
    super().__init__(num_features, eps, momentum, affine, True)
    self.process_group = process_group

  def _check_input_dim(self, x):
    if x.dim() < 2:
      raise ValueError('expected at least 2D input (got {}D input)'.format(x.dim()))

  def forward(self, x):
    self._check_input_dim(x)
    if not self.training:
      return nn.functional.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, False, 0.0,
                                      self.eps)
    squeeze = x.dim() == 2
    if squeeze:
      x = x.unsqueeze(-1)
    out, mean, var, count = _SyncBatchNormFn.apply(x, self.weight, self.bias, self.eps, self.process_group)
    with torch.no_grad():
      self.num_batches_tracked += 1
      unbiased = var * (count / (count - 1).clamp_min(1))
      self.running_mean.mul_(1 - self.momentum).add_(mean.to(self.running_mean.dtype), alpha=self.momentum)
      self.running_var.mul_(1 - self.momentum).add_(unbiased.to(self.running_var.dtype), alpha=self.momentum)
    return out.squeeze(-1) if squeeze else out


SynchronizedBatchNorm1d = SynchronizedBatchNorm2d = SynchronizedBatchNorm3d = SynchronizedBatchNorm


def convert_model(module, process_group=None):
  """Replace every BatchNorm{1,2,3}d in `module` (recursively) by a SynchronizedBatchNorm carrying the same
  parameters and running statistics; reference lib/nn/sync_batchnorm/batchnorm.py:353-393."""
  mod = module
  if isinstance(module, nn.modules.batchnorm._BatchNorm) and not isinstance(module, SynchronizedBatchNorm):
    mod = SynchronizedBatchNorm(module.num_features, module.eps, module.momentum, module.affine, process_group)
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
