"""Micro-benchmark of the fused DiffNet layer-stack kernel: CUDA-event time per 20-layer launch at cfg2-like sizes.
PTTSPP_DIFFNET_DBG selects timing experiments (see csrc/diffnet_layer.h); run one variant per process."""
import math
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from promptttspp_b200 import ops  # noqa: E402
from test_gpu_diffnet import _make_layers  # noqa: E402

torch.set_grad_enabled(False)
C_ = 256


def main():
    B = int(os.environ.get("BD_B", 16))
    T = int(os.environ.get("BD_T", 2582))
    NL = int(os.environ.get("BD_L", 20))
    reps = int(os.environ.get("BD_REPS", 20))
    layers = _make_layers(NL, 1)
    stack = ops.DiffNetStack(layers, "cuda")
    g = torch.Generator().manual_seed(2)
    h = torch.randn(B, T, C_, generator=g).cuda()
    cond = (torch.randn(NL, B, T, 2 * C_, generator=g) * 0.5).cuda()
    step = (torch.randn(NL + 1, C_, generator=g) * 0.3).cuda()
    y0 = ops.split_f16(h, step[0].contiguous())
    y1 = tuple(torch.zeros_like(t) for t in y0)
    skip = torch.zeros(B, T, C_, device="cuda")
    sp = tuple(torch.zeros_like(t) for t in y0)
    done = stack.new_flags(B, T)
    epoch = 0

    def run():
        nonlocal epoch
        epoch += 1
        stack.run(cond, step, (y0, y1), skip, done, epoch=epoch, skip_planes=sp)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rows = B * T
    fl = rows * 2.0 * (3 * C_ * 2 * C_ + C_ * 2 * C_) * NL
    n_mt = math.ceil(T / 128)
    units = math.ceil(B * n_mt / 2)
    if os.environ.get("BD_PROF"):
        prof = torch.zeros(74, 16, dtype=torch.int64, device="cuda")
        epoch += 1
        stack.run(cond, step, (y0, y1), skip, done, epoch=epoch, skip_planes=sp, dbg_prof=prof)
        torch.cuda.synchronize()
        p = prof.double().cpu()
        tot = p[:, 7].clamp_min(1)
        names = {0: "MMA wait tempty (D)", 1: "MMA wait tempty (O)", 2: "MMA wait fullA", 3: "MMA wait fullB (D)",
                 4: "MMA wait fullB (O)", 5: "MMA wait zfull", 8: "TMA wait emptyA", 9: "TMA wait emptyB",
                 10: "TMA wait prev-layer flags", 11: "EPI w0 wait tfull (D)", 12: "EPI w0 wait tfull (O)"}
        print(f"   MMA-warp loop: {float(tot.mean()) / 1e3:.0f} kcycles per cluster (mean), fractions of it:")
        for k, n in names.items():
            fr = p[:, k] / tot
            print(f"     {n:28s} mean {float(fr.mean()) * 100:5.1f}%  max {float(fr.max()) * 100:5.1f}%")
    print(f"dbg={os.environ.get('PTTSPP_DIFFNET_DBG', '0')} B={B} T={T} L={NL}: {ms * 1e3 / NL:8.1f} us/layer  {ms:7.3f} ms/launch  "
          f"{fl / ms / 1e9:7.1f} TFLOP/s algorithmic ({3 * fl / ms / 1e9:7.1f} issued)  units {units} -> {units * NL / 74:.1f} tasks/cluster",
          flush=True)


if __name__ == "__main__":
    main()
