"""One forward + backward of the attention core at the C = 256 hierarchy shape (for ncu):
    ncu --set full --clock-control none --import-source on -k regex:attn_ -o gpurun_out/attn python tools/profile_attention_bwd.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsg_b200.models.heads.transformer import attention_core

dev = torch.device('cuda:0')
b, h, l, s, hd = 256, 4, 256, 256, 64
q, k, v = [torch.randn(b * h, n, hd, device=dev, requires_grad=True) for n in (l, s, s)]
mask = torch.zeros(b, s, dtype=torch.bool, device=dev)
mask[:, int(0.8 * s):] = True
w = torch.randn(b * h, l, hd, device=dev)
(attention_core(q, k, v, mask, b, h) * w).sum().backward()
torch.cuda.synchronize()
print('ok')
