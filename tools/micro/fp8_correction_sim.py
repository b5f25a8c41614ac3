"""CPU simulation (no GPU): accuracy of the NCE similarities if the two cross terms of the three-pass fp16 split
(el.ph + eh.pl) ran as ONE e4m3 pass over concatenated operands instead of two fp16 passes.
Accumulation is idealised as exact; torch's float8_e4m3fn does the operand rounding.
    python tools/micro/fp8_correction_sim.py"""
import math
import torch

torch.manual_seed(0)
N, P, D = 1024, 2048, 256
e = torch.randn(N, D, dtype=torch.float64); e /= e.norm(dim=1, keepdim=True)
p = torch.randn(P, D, dtype=torch.float64); p /= p.norm(dim=1, keepdim=True)
sc = math.sqrt(16 * math.log2(math.e))           # operands pre-scaled: the accumulator is the base-2 exponent
es, ps = e.float() * sc, p.float() * sc
exact = es.double() @ ps.double().T


def split(x):
  hi = x.half()
  return hi, (x - hi.float()).half()


def mm(a, b):
  return a.double() @ b.double().T


def q8(x):
  return x.float().to(torch.float8_e4m3fn).float()


eh, el = split(es)
ph, pl = split(ps)
s3 = mm(eh, ph) + mm(el, ph) + mm(eh, pl)
k = 2.0 ** 11
s8 = mm(eh, ph) + (mm(q8(el.float() * k), q8(ph.float())) + mm(q8(eh.float()), q8(pl.float() * k))) / k
lse = lambda s: torch.logsumexp(s * math.log(2), dim=1)
for name, s in (('three fp16 passes', s3), ('fp16 + one e4m3 correction pass', s8), ('one fp16 pass', mm(eh, ph)),
                ('plain fp32 GEMM', (es @ ps.T).double())):
  err = (s - exact).abs()
  print('%-34s max |err| of the exponent %.2e  rms %.2e   log-sum-exp |err| %.2e (value %.2f)' % (
      name, err.max().item(), (s - exact).pow(2).mean().sqrt().item(), (lse(s) - lse(exact)).abs().max().item(),
      lse(exact).mean().item()))
