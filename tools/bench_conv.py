"""Micro-benchmark of the conv kernels on DiffNet / BigVGAN shapes (CUDA-event timing, algorithmic TFLOP/s)."""
import math
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from promptttspp_b200 import _abi, ops  # noqa: E402

torch.set_grad_enabled(False)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B, T, C = 16, 2582, 256  # cfg2: 41 312 padded frames
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, C, generator=g).cuda()
    planes = ops.split_f16(x)
    cond = torch.randn(B, T, 2 * C, generator=g).cuda()
    w1 = torch.randn(2 * C, C, 3, generator=g) / math.sqrt(3 * C)
    w2 = torch.randn(C, C, 1, generator=g) / math.sqrt(C)
    b1 = torch.randn(2 * C, generator=g).cuda()
    b2 = torch.randn(C, generator=g).cuda()
    w1s = ops.pack_conv_weight_split(w1, interleave_halves=True, device="cuda")
    w2s = ops.pack_conv_weight_split(w2, device="cuda")
    w1p = ops.pack_conv_weight(w1, interleave_halves=True, device="cuda")
    w2p = ops.pack_conv_weight(w2, device="cuda")
    h = x.clone()
    skip = torch.zeros_like(x)
    out512 = torch.zeros(B, T, 2 * C, device="cuda")
    fl1 = 2.0 * B * T * 2 * C * C * 3
    fl2 = 2.0 * B * T * C * C
    rows = []

    def pv(order=None, ew=None):
        e = {"PTTSPP_UMMA_PAIR": "2"}
        if order is not None:
            e["PTTSPP_UMMA_ORDER"] = str(order)
        if ew is not None:
            e["PTTSPP_UMMA_EW"] = str(ew)
        return e

    # experiment bits: 2 one MMA per product, 4 no operand loads, 8 no stores, 32 no TMEM reads, 64 no epilogue
    VARIANTS = (("str", {"PTTSPP_UMMA_PAIR": "0", "PTTSPP_UMMA_EPI": "co"}), ("tma", {"PTTSPP_UMMA_PAIR": "0"}),
                ("pair", pv()), ("p-noLD", pv(4)), ("p-noST", pv(8)), ("p-noLS", pv(12)), ("p-noEpi", pv(64)),
                ("p-1mma", pv(2)))
    KEYS = ("PTTSPP_UMMA_PAIR", "PTTSPP_UMMA_EPI", "PTTSPP_UMMA_AS", "PTTSPP_UMMA_ORDER", "PTTSPP_UMMA_EW", "PTTSPP_UMMA_RL", "PTTSPP_UMMA_NACC")
    if os.environ.get("BENCH_QUICK"):
        VARIANTS = VARIANTS[1:]

    def add(name, fl, fn):
        modes = VARIANTS if name.startswith("umma") else (("    ", {}),)
        for tag, env in modes:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            _abi.lib().pttspp_debug_reload_env()
            ms = timeit(fn)
            rows.append((name, ms, fl / ms / 1e9))
            print(f"{tag:8s} {name:58s} {ms*1e3:9.1f} us   {fl / ms / 1e9:8.1f} TFLOP/s (algorithmic)", flush=True)
        for k in KEYS:
            os.environ.pop(k, None)
        _abi.lib().pttspp_debug_reload_env()

    for dil in (1, 8):
        add(f"umma dilated k3 d{dil} 256->512 plain fp32 out", fl1,
            lambda: ops.conv1d_umma_cl(planes, w1s, 2 * C, bias=b1, K=3, dil=dil, pad=dil, out=out512))
        add(f"umma dilated k3 d{dil} 256->512 gate+addend -> planes only", fl1,
            lambda: ops.conv1d_umma_cl(planes, w1s, 2 * C, bias=b1, K=3, dil=dil, pad=dil, act=ops.ACT_GATE, addend=cond,
                                       emit_planes=True, write_f32=False))
    add("umma 1x1 256->256 residual in place + planes", fl2,
        lambda: ops.conv1d_umma_cl(planes, w2s, C, bias=b2, res=h, out=h, out_div=math.sqrt(2.0), emit_planes=True,
                                   plane_add=b2))
    add("umma 1x1 256->256 skip accumulate", fl2,
        lambda: ops.conv1d_umma_cl(planes, w2s, C, bias=b2, out=skip, beta=1.0))
    w2full = ops.pack_conv_weight_split(torch.randn(2 * C, C, 1, generator=g) / math.sqrt(C), device="cuda")
    b2f = torch.randn(2 * C, generator=g).cuda()
    add("umma 1x1 256->512 DUAL residual+planes | skip accumulate", 2 * fl2,
        lambda: ops.conv1d_umma_dual_cl(planes, w2full, C,
                                        dict(bias=b2f[:C], res=h, out=h, out_div=math.sqrt(2.0), emit_planes=True,
                                             plane_add=b2),
                                        dict(bias=b2f[C:], out=skip, beta=1.0)))
    if os.environ.get("BENCH_QUICK"):
        return
    add("simt dilated k3 d1 256->512 gate+addend", fl1,
        lambda: ops.conv1d_cl(x, w1p, 2 * C, bias=b1, K=3, dil=1, pad=1, act=ops.ACT_GATE, addend=cond, in_add=b2, impl=1))
    add("simt 1x1 256->256 residual", fl2, lambda: ops.conv1d_cl(x, w2p, C, bias=b2, res=h, out=h, out_div=1.41421, impl=1))
    # BigVGAN-like: C=128/64/32 with long T
    for Cc, L, k, d in ((128, 30720, 7, 1), (64, 122880, 11, 5), (32, 245760, 3, 1)):
        xx = torch.randn(4, L, Cc, generator=torch.Generator().manual_seed(1)).cuda()
        ww = ops.pack_conv_weight(torch.randn(Cc, Cc, k) / math.sqrt(Cc * k), device="cuda")
        add(f"simt C={Cc} k{k} d{d} L={L} B=4", 2.0 * 4 * L * Cc * Cc * k,
            lambda: ops.conv1d_cl(xx, ww, Cc, K=k, dil=d, pad=(k * d - d) // 2, impl=1))
        if Cc % 64 == 0:
            pp = ops.split_f16(xx)
            ws = ops.pack_conv_weight_split(torch.randn(Cc, Cc, k) / math.sqrt(Cc * k), device="cuda")
            add(f"umma C={Cc} k{k} d{d} L={L} B=4", 2.0 * 4 * L * Cc * Cc * k,
                lambda: ops.conv1d_umma_cl(pp, ws, Cc, K=k, dil=d, pad=(k * d - d) // 2))
        from promptttspp_b200.layers.activations import AntiAliasActivation
        act = AntiAliasActivation(Cc)
        al = torch.zeros(Cc).cuda()
        uf, df = act.up.filter.view(-1).cuda(), act.down.lowpass.filter.view(-1).cuda()
        ms = timeit(lambda: ops.aa_snake_cl(xx, al, uf, df))
        print(f"aa_snake C={Cc} L={L} B=4: {ms*1e3:9.1f} us  {2 * 4 * xx.numel() / ms / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
