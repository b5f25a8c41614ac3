"""BASELINE.json configs[4]: flat spherical k-means sweep on one GPU.
    python tools/sweep_kmeans.py [out.txt]
X = normalise(N(0,1)) [N,D] fp32, init = randint(0,K) (seed 235), T = 20 iterations through the
reference-facing operator (kmeans_with_initial_labels).  Per point: ms per iteration, algorithmic
HBM rate N*(4D+8) bytes per iteration against the measured copy peak, and the dot-product rate
2*N*D*K flop per iteration.  Points whose CUDA-core E-step (K > 256 or D > 256: no tensor-core
path yet) would need more than ~5e13 flop in total are skipped and listed."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsg_b200 import ops
from hsg_b200.utils.segsort import common as S

T = 20
dev = torch.device('cuda:0')
peak, tpeak = 6539.2, 1357.1
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.exists(pk):
  peak = json.load(open(pk))['hbm_gbs']
  tpeak = json.load(open(pk)).get('bf16_tflops_sustained', tpeak)
# which roof bounds a point: one fp16 pass costs 2NDK flop against N(4D+8) bytes, i.e. the tensor roof is the lower one
# when K > peak_flops / peak_bytes * (4D+8) / (2D) ~ 425 at D=256; both fractions are printed
lines = ['# flat spherical k-means sweep (BASELINE configs[4]), 1 x B200, T=%d, HBM peak %.0f GB/s (measured copy), tensor peak %.0f TFLOP/s (measured cuBLAS bf16, sustained)' % (T, peak, tpeak),
         '%9s %4s %5s %6s %10s %10s %8s %10s %8s %7s' % ('N', 'D', 'K', 'E-step', 'ms/iter', 'GB/s(alg)', 'hbm', 'TFLOP/s', 'tensor', 'bound')]
for nn in (100000, 1000000, 10000000):
  for d in (64, 256, 512):
    for k in (32, 256, 2048):
      tc = bool(ops.tc_d16(d, k))
      flops = 2.0 * nn * d * k
      if not tc and flops * T > 5e13:
        lines.append('%9d %4d %5d %6s %10s' % (nn, d, k, 'simt', 'skipped'))
        continue
      g = torch.Generator(device=dev); g.manual_seed(235)
      x = torch.randn(nn, d, device=dev, generator=g)
      x = x / x.norm(dim=1, keepdim=True)
      init = torch.randint(0, k, (nn,), device=dev, generator=g)
      S.kmeans_with_initial_labels(x, init, k, 2)
      torch.cuda.synchronize()
      t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      t0.record()
      S.kmeans_with_initial_labels(x, init, k, T)
      t1.record(); torch.cuda.synchronize()
      ms = t0.elapsed_time(t1) / T
      gbs = nn * (4.0 * d + 8) / ms / 1e6
      tfs = flops / ms / 1e9
      bound = 'tensor' if tfs / tpeak > gbs / peak else 'hbm'
      lines.append('%9d %4d %5d %6s %10.3f %10.1f %8.3f %10.1f %8.3f %7s' % (nn, d, k, 'tc' if tc and nn >= 16384 else 'simt', ms, gbs,
                                                                            gbs / peak, tfs, tfs / tpeak, bound))
      del x, init
      torch.cuda.empty_cache()
text = '\n'.join(lines)
print(text)
if len(sys.argv) > 1:
  open(sys.argv[1], 'w').write(text + '\n')
