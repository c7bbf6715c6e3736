"""Is the AA-Snake pair kernel bound by the SM or by the memory system?  The same kernel on an L2-resident problem
(B=1: 31 MB in + 31 MB out at the 64-channel stage) against the HBM-resident one (B=16)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from promptttspp_b200 import ops  # noqa: E402
from promptttspp_b200.layers.activations import AntiAliasActivation  # noqa: E402

torch.set_grad_enabled(False)
act = AntiAliasActivation(4)
up, down = act.up.filter.view(-1).cuda(), act.down.lowpass.filter.view(-1).cuda()
for C, L in ((64, 122880), (32, 245760)):
    for B in (1, 2, 4, 16):
        x = torch.randn(B, L, C, device="cuda")
        alpha = (torch.rand(C, device="cuda") - 0.5)
        for _ in range(3):
            ops.aa_snake_cl(x, alpha, up, down, pair=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            ops.aa_snake_cl(x, alpha, up, down, pair=True)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / n
        print(f"C={C:3d} B={B:2d}: {us:8.1f} us  {x.numel() / us / 1e3:6.2f} Gelem/s  ({2 * 4 * x.numel() / 1e6:7.1f} MB in+out)")
