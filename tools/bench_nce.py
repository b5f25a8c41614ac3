"""Micro-benchmark of the NCE forward (tensor-core kernel against the fp32 CUDA-core
kernel on a sample): python tools/bench_nce.py [N] [P] [D]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hsg_b200
from hsg_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 48 * 448 * 448
P = int(sys.argv[2]) if len(sys.argv) > 2 else 12288
D = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device('cuda:0')
lib = hsg_b200.load_library()
g = torch.Generator(device=dev); g.manual_seed(235)
e = torch.randn(N, D, device=dev, generator=g); e = e / e.norm(dim=1, keepdim=True)
pr = torch.randn(P, D, device=dev, generator=g); pr = pr / pr.norm(dim=1, keepdim=True)
inst = torch.randint(0, P, (N,), device=dev, generator=g)
psem = torch.stack([torch.arange(P, device=dev) // 256, torch.arange(P, device=dev)], 0)
sem = torch.stack([psem[0][inst], inst], 0)
sets = ['segsort+', 'segsort+']
for _ in range(2):
  ll = ops.nce_log_likelihood(e, inst, sem, pr, psem, 16.0, sets)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
reps = 3
for _ in range(reps):
  ll = ops.nce_log_likelihood(e, inst, sem, pr, psem, 16.0, sets)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / reps
print('tensor cores: %.2f ms  %.1f TFLOP/s (3 fp16 passes)  mean loss %s' % (ms, 2.0 * N * P * 3 * D / ms / 1e9, ll.mean(1).tolist()))
n_s = min(N, 65536)
lib.hsg_debug_set_flags(1)        # fp32 CUDA-core kernel
ref = ops.nce_log_likelihood(e[:n_s], inst[:n_s], sem[:, :n_s].contiguous(), pr, psem, 16.0, sets)
lib.hsg_debug_set_flags(0)
rel = ((ll[:, :n_s] - ref).abs() / ref.abs().clamp_min(1e-3)).max().item()
print('max relative difference to the fp32 kernel on %d pixels: %.3g' % (n_s, rel))
assert rel < 2e-5

# backward (dE, dP): recomputed G chunk + two fp32-grade GEMMs
nb = min(N, 1 << 20)
eb = e[:nb].clone().requires_grad_(True)
pb = pr.clone().requires_grad_(True)
semb = sem[:, :nb].contiguous()
for flags, name in ((0, 'G chunk and GEMMs on tensor cores'), (8, 'G chunk on CUDA cores, GEMMs on tensor cores'),
                    (4, 'all on CUDA cores')):
  lib.hsg_debug_set_flags(flags)
  grads = []
  for rep in range(2):
    eb.grad = None; pb.grad = None
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ll = ops.nce_log_likelihood(eb, inst[:nb], semb, pb, psem, 16.0, sets)
    loss = ll.mean()
    t0.record()
    loss.backward()
    t1.record(); torch.cuda.synchronize()
  print('backward on %d pixels, %s: %.1f ms' % (nb, name, t0.elapsed_time(t1)))
  grads.append((eb.grad.clone(), pb.grad.clone()))
  if flags == 0:
    ref_g = grads[-1]
  else:
    for a_, b_ in zip(ref_g, grads[-1]):
      rel = ((a_ - b_).norm() / b_.norm()).item()
      print('  relative difference to the tensor-core gradients: %.3g' % rel)
      assert rel < 1e-5
lib.hsg_debug_set_flags(0)
