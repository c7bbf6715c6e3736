"""AA-Snake micro-benchmark at the four BigVGAN stage shapes of cfg3 (B=16 x 1024 frames): strip kernel vs channel-pair kernel."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from promptttspp_b200 import ops  # noqa: E402
from promptttspp_b200.layers.activations import AntiAliasActivation  # noqa: E402

torch.set_grad_enabled(False)
act = AntiAliasActivation(4)
up, down = act.up.filter.view(-1).cuda(), act.down.lowpass.filter.view(-1).cuda()
for C, L in ((256, 6144), (128, 30720), (64, 122880), (32, 245760)):
    x = torch.randn(16, L, C, device="cuda")
    alpha = (torch.rand(C, device="cuda") - 0.5)
    for pair in (False, True):
        for _ in range(2):
            ops.aa_snake_cl(x, alpha, up, down, pair=pair)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.aa_snake_cl(x, alpha, up, down, pair=pair)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        gb = 2 * 4 * x.numel() / 1e9
        print(f"C={C:4d} L={L:7d} {'pair ' if pair else 'strip'}: {us:8.1f} us  {gb / us * 1e6:7.1f} GB/s algorithmic (fp32 in, fp32 out)")
