"""Clustering-transformer attention (K5) at the hierarchy's shapes (BASELINE configs[3]) and a batch sweep:
our fused kernel (fwd, fwd+bwd) against torch's nn.MultiheadAttention core as the reference runs it
(need_weights=True slow path: scale, baddbmm with the -inf mask, softmax, bmm; fp32, TF32 off).
    python tools/bench_attention.py [out.txt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsg_b200 import _lib
from hsg_b200.models.heads.transformer import attention_core

lib = _lib.load()

dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False


def torch_core(q, k, v, mask, b, h):
  hd = q.shape[-1]
  add = torch.zeros(mask.shape, dtype=q.dtype, device=q.device).masked_fill(mask, float('-inf'))
  add = add.view(b, 1, 1, -1).expand(b, h, 1, mask.shape[1]).reshape(b * h, 1, -1)
  sc = torch.baddbmm(add, q / hd ** 0.5, k.transpose(1, 2))
  return torch.bmm(torch.softmax(sc, -1), v)


def timeit(fn, reps=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record()
  for _ in range(reps):
    fn()
  t1.record(); torch.cuda.synchronize()
  return t0.elapsed_time(t1) / reps * 1e3


lines = ['# attention core, fp32, 4 heads x hd 64 (C=256); us per call; flop = 4*B*h*L*S*hd (fwd)',
         '# fwd+bwd columns: default dispatch | backward forced to the CUDA-core kernels | backward forced to tcgen05 | torch',
         '# fwd columns: default dispatch | forced to the CUDA-core kernel | forced to tcgen05 | torch',
         '%5s %4s %4s %10s %10s %10s %10s %9s %14s %12s %12s %14s %10s' % ('B', 'L', 'S', 'ours fwd', 'simt fwd', 'tc fwd', 'torch fwd', 'speed-up',
                                                                          'ours fwd+bwd', 'simt bwd', 'tc bwd', 'torch fwd+bwd', 'GFLOP/s')]
for b in (8, 16, 64, 256, 1024):
  for (l, s) in ((256, 256), (64, 256), (16, 64)):
    h, hd = 4, 64
    q = torch.randn(b * h, l, hd, device=dev, requires_grad=True)
    k = torch.randn(b * h, s, hd, device=dev, requires_grad=True)
    v = torch.randn(b * h, s, hd, device=dev, requires_grad=True)
    mask = torch.zeros(b, s, dtype=torch.bool, device=dev)
    mask[:, int(0.8 * s):] = True
    w = torch.randn(b * h, l, hd, device=dev)
    with torch.no_grad():
      f_ours = timeit(lambda: attention_core(q, k, v, mask, b, h))
      lib.hsg_debug_set_flags(16)
      f_simt = timeit(lambda: attention_core(q, k, v, mask, b, h))
      lib.hsg_debug_set_flags(32)
      f_tc = timeit(lambda: attention_core(q, k, v, mask, b, h))
      lib.hsg_debug_set_flags(0)
      f_torch = timeit(lambda: torch_core(q, k, v, mask, b, h))

    def fb(fn):
      def run():
        for t_ in (q, k, v):
          t_.grad = None
        (fn(q, k, v, mask, b, h) * w).sum().backward()
      return run
    fb_ours = timeit(fb(attention_core))
    lib.hsg_debug_set_flags(64)
    fb_simt = timeit(fb(attention_core))
    lib.hsg_debug_set_flags(128)
    fb_tc = timeit(fb(attention_core))
    lib.hsg_debug_set_flags(0)
    fb_torch = timeit(fb(torch_core))
    lines.append('%5d %4d %4d %10.1f %10.1f %10.1f %10.1f %9.2f %14.1f %12.1f %12.1f %14.1f %10.0f' % (
        b, l, s, f_ours, f_simt, f_tc, f_torch, f_torch / f_ours, fb_ours, fb_simt, fb_tc, fb_torch,
        4.0 * b * h * l * s * hd / f_ours / 1e3))
text = '\n'.join(lines)
print(text)
if len(sys.argv) > 1:
  open(sys.argv[1], 'w').write(text + '\n')
