"""Timing of the encoder's rel-pos attention at cfg2's shape (B=16, T=256, 2 heads x 128): the fused tcgen05 kernel
(csrc/attention_umma.cu) or, with PTTSPP_ATTN_TC=0, the two CUDA-core launches (csrc/attention.cu).  Also the target
of the `ncu --set full` capture under profiles/."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from promptttspp_b200 import ops  # noqa: E402

torch.set_grad_enabled(False)


def main():
    B, T, H, dk = int(os.environ.get("BA_B", 16)), int(os.environ.get("BA_T", 256)), 2, 128
    legacy = os.environ.get("BA_LEGACY", "1") == "1"
    reps = int(os.environ.get("BA_REPS", 50))
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(B, T, H * dk, generator=g).cuda() for _ in range(3))
    Tp = T if legacy else 2 * T - 1
    p = torch.randn(Tp, H * dk, generator=g).cuda()
    bu, bv = (torch.randn(H, dk, generator=g) * 0.3).cuda(), (torch.randn(H, dk, generator=g) * 0.3).cuda()
    lens = torch.randint(T // 2, T + 1, (B,), generator=g).cuda()
    for _ in range(3):
        ops.relpos_attention(q, k, v, p, bu, bv, lens, H, legacy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.relpos_attention(q, k, v, p, bu, bv, lens, H, legacy)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    flops = 2.0 * B * H * T * dk * (2.0 * T + Tp)
    print(f"attention B={B} T={T} legacy={legacy} tc={os.environ.get('PTTSPP_ATTN_TC', '1')}: {us:8.1f} us/call  "
          f"{flops / us / 1e6:7.2f} TFLOP/s algorithmic ({3 * flops / us / 1e6:.2f} issued split-fp16)")


if __name__ == "__main__":
    main()
