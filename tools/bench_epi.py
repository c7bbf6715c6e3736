"""Isolate the per-tile epilogue cost of the tcgen05 conv: 1x1 convs with a single K slab (Cin=64)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from promptttspp_b200 import ops  # noqa: E402
from tools.bench_conv import timeit  # noqa: E402

torch.set_grad_enabled(False)
B, T = 16, 2048
g = torch.Generator().manual_seed(0)
for Cin, Cout in ((64, 128), (64, 256), (64, 512), (256, 256), (256, 512)):
    x = torch.randn(B, T, Cin, generator=g).cuda()
    planes = ops.split_f16(x)
    w = ops.pack_conv_weight_split(torch.randn(Cout, Cin, 1, generator=g) / math.sqrt(Cin), device="cuda")
    out = torch.zeros(B, T, Cout, device="cuda")
    res = torch.randn(B, T, Cout, generator=g).cuda()
    tiles = B * T // 128 * (Cout // 128)
    for name, fn in (
        ("fp32 out", lambda: ops.conv1d_umma_cl(planes, w, Cout, out=out)),
        ("planes only", lambda: ops.conv1d_umma_cl(planes, w, Cout, emit_planes=True, write_f32=False)),
        ("accumulate (RMW)", lambda: ops.conv1d_umma_cl(planes, w, Cout, out=out, beta=1.0)),
        ("residual in place", lambda: ops.conv1d_umma_cl(planes, w, Cout, res=res, out=res, out_div=1.4142135)),
    ):
        ms = timeit(fn)
        per_tile = ms * 1e3 / (tiles / 148.0)
        print(f"Cin={Cin:4d} Cout={Cout:4d} {name:20s} {ms*1e3:8.1f} us  tiles/SM {tiles/148:5.2f}  -> {per_tile:6.2f} us per tile-slot",
              flush=True)
