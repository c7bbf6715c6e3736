"""Tiny driver for ncu: the two tcgen05 launches of a DiffNet layer on the cfg2 shape."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from promptttspp_b200 import ops  # noqa: E402

torch.set_grad_enabled(False)
B, T, C = 16, 2048, 256
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, C, generator=g).cuda()
planes = ops.split_f16(x)
cond = torch.randn(B, T, 2 * C, generator=g).cuda()
w1 = torch.randn(2 * C, C, 3, generator=g) / math.sqrt(3 * C)
b1 = torch.randn(2 * C, generator=g).cuda()
w1s = ops.pack_conv_weight_split(w1, interleave_halves=True, device="cuda")
w2s = ops.pack_conv_weight_split(torch.randn(2 * C, C, 1, generator=g) / 16, device="cuda")
b2 = torch.randn(2 * C, generator=g).cuda()
h = x.clone()
skip = torch.zeros_like(x)
for _ in range(3):
    _, zp = ops.conv1d_umma_cl(planes, w1s, 2 * C, bias=b1, K=3, dil=2, pad=2, act=ops.ACT_GATE, addend=cond,
                               emit_planes=True, write_f32=False)
    ops.conv1d_umma_dual_cl(zp, w2s, C,
                            dict(bias=b2[:C], res=h, out=h, out_div=math.sqrt(2.0), emit_planes=True, plane_add=b1[:C]),
                            dict(bias=b2[C:], out=skip, beta=1.0))
torch.cuda.synchronize()
