"""GPU: the tcgen05/TMA split-fp16 conv path against torch fp32 conv1d on the CPU."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def ops():
    from promptttspp_b200 import _abi, ops

    _abi.check(_abi.lib().pttspp_device_check())
    return ops


def _cl(x):
    return x.transpose(1, 2).contiguous().cuda()


def _bct(x):
    return x.float().cpu().transpose(1, 2)


def test_split_planes_reconstruct(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 50, 64, generator=g) * 3
    add = torch.randn(64, generator=g)
    hi, lo = ops.split_f16(x.cuda(), add.cuda())
    rec = hi.float().cpu() + lo.float().cpu()
    ref = x + add
    # 2^-22 relative while lo is a normal fp16; 2^-25 absolute floor where lo goes subnormal
    assert float(((rec - ref).abs() - 4e-7 * ref.abs()).max()) < 4e-8


def test_split_planes_saturate_instead_of_overflowing(ops):
    """|v| beyond fp16's range: the hi plane saturates at 65504 (no inf), the lo plane carries the rest up to 2 x 65504;
    nothing downstream can become inf - inf = NaN (activations had no saturation)."""
    x = torch.tensor([[1.0e5, -1.0e5, 3.0e5, -7.0e8, 65504.0, 65519.9, 65520.0, 1.0]] * 2).view(1, 2, 8)
    x = torch.cat([x, torch.zeros(1, 2, 8)], dim=2).contiguous()  # 16 channels
    hi, lo = ops.split_f16(x.cuda())
    hi, lo = hi.float().cpu(), lo.float().cpu()
    assert torch.isfinite(hi).all() and torch.isfinite(lo).all()
    rec = (hi + lo)[0, 0, :8]
    assert rec[0] == 1.0e5 and rec[1] == -1.0e5          # exact through the lo plane
    assert rec[2] == 2 * 65504.0 and rec[3] == -2 * 65504.0  # clipped, finite
    assert rec[4] == 65504.0 and abs(float(rec[5]) - 65519.9) < 0.01 and rec[6] == 65520.0 and rec[7] == 1.0


UMMA_CASES = [
    # Cin, Cout, K, dil, B, T
    (64, 128, 1, 1, 1, 128),      # single stage, single tile
    (256, 512, 3, 1, 2, 300),     # DiffNet dilated conv shapes
    (256, 512, 3, 8, 2, 259),
    (256, 256, 1, 1, 3, 77),
    (256, 80, 1, 1, 2, 130),      # Cout < BN (TMA zero-fills the missing weight rows)
    (128, 128, 7, 1, 1, 129),
    (1024, 256, 9, 1, 1, 200),    # many K iterations through the 3-stage ring
    (64, 64, 11, 5, 1, 515),
    # enough 128-row units (>= half the SMs) to take the A-stationary kernel (taps share one halo tile)
    (256, 512, 3, 1, 4, 2500),
    (256, 512, 3, 2, 4, 2431),
    (256, 512, 3, 4, 3, 3333),
    (256, 512, 3, 8, 4, 2500),
    (256, 256, 1, 1, 4, 2500),
    (256, 80, 1, 1, 4, 2500),
    (128, 128, 7, 1, 2, 6000),
    (128, 128, 11, 5, 2, 6000),
    (64, 64, 11, 5, 1, 12000),
    # >= 2 units per SM: the A-stationary 64-column kernel with 256-row units (two MMA tiles share every weight stage);
    # odd number of 128-row tiles per utterance
    (64, 64, 11, 5, 4, 12100),
    (64, 64, 11, 1, 4, 12100),
    (256, 256, 17, 1, 4, 2500),   # frame-prior shape: halo of 16 rows
    (256, 256, 5, 1, 4, 2500),
    # 32 channels: weight-resident SWIZZLE_64B kernel (BigVGAN's last stage)
    (32, 32, 3, 1, 2, 1000),
    (32, 32, 3, 5, 2, 3001),
    (32, 32, 7, 3, 1, 5000),
    (32, 32, 11, 1, 3, 2999),
    (32, 32, 11, 5, 2, 40000),
    (32, 32, 1, 1, 1, 77),
    # 64 channels, <= 7 taps: the same weight-resident kernel with SWIZZLE_128B rows
    (64, 64, 3, 1, 2, 1000),
    (64, 64, 3, 5, 2, 3001),
    (64, 64, 7, 5, 2, 20000),
    (64, 64, 7, 1, 3, 2999),
]


VARIANTS = {
    # env of the kernel selection in conv1d_umma_launch
    "stream": {"PTTSPP_UMMA_PAIR": "0", "PTTSPP_UMMA_EPI": "co"},      # streaming kernel, coalescing epilogue
    "stream_tma": {"PTTSPP_UMMA_PAIR": "0"},                           # ... TMA bulk-store epilogue (its default)
    "astat": {"PTTSPP_UMMA_PAIR": "0", "PTTSPP_UMMA_AS": "1"},         # single-CTA A-stationary (opt-in)
    "pair": {"PTTSPP_UMMA_PAIR": "2"},                                 # CTA-pair kernel forced wherever it applies
    # ... gate epilogue through the staging tile, two accumulators even for short contractions
    "pair_co": {"PTTSPP_UMMA_PAIR": "2", "PTTSPP_UMMA_RL": "0", "PTTSPP_UMMA_NACC": "2"},
}


@pytest.fixture(autouse=True)
def _clean_knobs():
    """set up before (and torn down after) monkeypatch: the library's cached knobs follow the restored environment"""
    yield
    from promptttspp_b200 import _abi

    _abi.lib().pttspp_debug_reload_env()


def _select(monkeypatch, variant):
    from promptttspp_b200 import _abi

    for k in ("PTTSPP_UMMA_PAIR", "PTTSPP_UMMA_EPI", "PTTSPP_UMMA_AS", "PTTSPP_UMMA_RL", "PTTSPP_UMMA_NACC"):
        monkeypatch.delenv(k, raising=False)
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    _abi.lib().pttspp_debug_reload_env()  # the knobs are cached by the library


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("Cin,Cout,K,dil,B,T", UMMA_CASES)
def test_conv1d_umma_plain(ops, monkeypatch, Cin, Cout, K, dil, B, T, variant):
    # every tcgen05 kernel variant: streaming, A-stationary with taps sharing one halo tile, cta_group::2 pair
    _select(monkeypatch, variant)
    g = torch.Generator().manual_seed(Cin + Cout * 3 + K)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = torch.randn(Cout, generator=g)
    pad = (K * dil - dil) // 2
    ref = F.conv1d(x, w, b, padding=pad, dilation=dil)
    planes = ops.split_f16(_cl(x))
    out, _ = ops.conv1d_umma_cl(planes, ops.pack_conv_weight_split(w, device="cuda"), Cout, bias=b.cuda(), K=K, dil=dil,
                                pad=pad)
    err = float((_bct(out) - ref).abs().max())
    print(f"umma {Cin}->{Cout} k{K} d{dil}: max-abs err {err:.3e}")
    # the tensor core truncates on accumulate: the bias grows linearly with the K*Cin/16 accumulations; the CTA-pair
    # kernel keeps main and cross terms in ONE accumulator up to K*Cin = 256 (768 for the gated DiffNet conv)
    slope = 1.8e-8 if variant == "pair" and K * Cin <= 256 else 6e-9
    assert err < 1e-5 + slope * K * Cin, err


@pytest.mark.parametrize("variant", ["stream", "stream_tma", "pair", "pair_co"])
def test_conv1d_umma_diffnet_chain(ops, monkeypatch, variant):
    """gate conv -> planes -> 1x1 residual/skip convs, the two tcgen05 launches of a DiffNet layer."""
    _select(monkeypatch, variant)
    g = torch.Generator().manual_seed(6)
    B, C, T, dil = 2, 256, 333, 4
    x = torch.randn(B, C, T, generator=g)
    step = torch.randn(C, generator=g)
    cond = torch.randn(B, 2 * C, T, generator=g)
    w1 = torch.randn(2 * C, C, 3, generator=g) / math.sqrt(C * 3)
    b1 = torch.randn(2 * C, generator=g)
    w2 = torch.randn(2 * C, C, 1, generator=g) / math.sqrt(C)
    b2 = torch.randn(2 * C, generator=g)
    skip_old = torch.randn(B, C, T, generator=g)
    next_step = torch.randn(C, generator=g)
    y = F.conv1d(x + step[None, :, None], w1, b1, padding=dil, dilation=dil) + cond
    gate, filt = torch.chunk(y, 2, dim=1)
    z_ref = torch.sigmoid(gate) * torch.tanh(filt)
    o = F.conv1d(z_ref, w2, b2)
    x_ref = (x + o[:, :C]) / math.sqrt(2.0)
    skip_ref = skip_old + o[:, C:]
    perm = torch.empty(2 * C, dtype=torch.long)
    perm[0::2] = torch.arange(C)
    perm[1::2] = torch.arange(C) + C
    planes = ops.split_f16(_cl(x), step.cuda())
    _, zp = ops.conv1d_umma_cl(planes, ops.pack_conv_weight_split(w1, interleave_halves=True, device="cuda"), 2 * C,
                               bias=b1[perm].cuda(), K=3, dil=dil, pad=dil, act=ops.ACT_GATE,
                               addend=_cl(cond[:, perm]), emit_planes=True, write_f32=False)
    z = zp[0].float() + zp[1].float()
    assert float((_bct(z) - z_ref).abs().max()) < 2e-5
    xbuf = _cl(x)
    w2s_res = ops.pack_conv_weight_split(w2[:C], device="cuda")
    w2s_skip = ops.pack_conv_weight_split(w2[C:], device="cuda")
    _, yp = ops.conv1d_umma_cl(zp, w2s_res, C, bias=b2[:C].cuda(), res=xbuf, out=xbuf, out_div=math.sqrt(2.0),
                               emit_planes=True, plane_add=next_step.cuda())
    assert float((_bct(xbuf) - x_ref).abs().max()) < 2e-5
    ynext = yp[0].float() + yp[1].float()
    assert float((_bct(ynext) - (x_ref + next_step[None, :, None])).abs().max()) < 2e-5
    sbuf = _cl(skip_old)
    ops.conv1d_umma_cl(zp, w2s_skip, C, bias=b2[C:].cuda(), out=sbuf, beta=1.0)
    assert float((_bct(sbuf) - skip_ref).abs().max()) < 2e-5


@pytest.mark.parametrize("variant", ["stream", "stream_tma", "pair", "pair_co"])
def test_conv1d_umma_dual_epilogue(ops, monkeypatch, variant):
    """residual | skip halves of the DiffNet output projection in ONE launch."""
    _select(monkeypatch, variant)
    g = torch.Generator().manual_seed(12)
    B, C, T = 4, 256, 2600
    z = torch.rand(B, C, T, generator=g) * 2 - 1
    x = torch.randn(B, C, T, generator=g)
    skip_old = torch.randn(B, C, T, generator=g)
    w2 = torch.randn(2 * C, C, 1, generator=g) / math.sqrt(C)
    b2 = torch.randn(2 * C, generator=g)
    nxt = torch.randn(C, generator=g)
    o = F.conv1d(z, w2, b2)
    x_ref = (x + o[:, :C]) / math.sqrt(2.0)
    skip_ref = skip_old + o[:, C:]
    zp = ops.split_f16(_cl(z))
    xbuf, sbuf = _cl(x), _cl(skip_old)
    (o1, p1), (o2, p2) = ops.conv1d_umma_dual_cl(
        zp, ops.pack_conv_weight_split(w2, device="cuda"), C,
        dict(bias=b2[:C].cuda(), res=xbuf, out=xbuf, out_div=math.sqrt(2.0), emit_planes=True, plane_add=nxt.cuda()),
        dict(bias=b2[C:].cuda(), out=sbuf, beta=1.0, emit_planes=True))
    assert float((_bct(xbuf) - x_ref).abs().max()) < 2e-5
    assert float((_bct(sbuf) - skip_ref).abs().max()) < 2e-5
    assert float((_bct(p1[0].float() + p1[1].float()) - (x_ref + nxt[None, :, None])).abs().max()) < 2e-5
    assert float((_bct(p2[0].float() + p2[1].float()) - skip_ref).abs().max()) < 2e-5


@pytest.mark.parametrize("variant", ["stream_tma", "pair", "pair_co"])
@pytest.mark.parametrize("dil,T", [(1, 2500), (8, 2582)])
def test_conv1d_umma_gate_large(ops, monkeypatch, variant, dil, T):
    """DiffNet gated conv at bench-like sizes: fp32 output AND operand planes, ragged out_len mask, plane_add."""
    _select(monkeypatch, variant)
    g = torch.Generator().manual_seed(40 + dil)
    B, C = 4, 256
    x = torch.randn(B, C, T, generator=g)
    cond = torch.randn(B, 2 * C, T, generator=g)
    w1 = torch.randn(2 * C, C, 3, generator=g) / math.sqrt(C * 3)
    b1 = torch.randn(2 * C, generator=g)
    padd = torch.randn(C, generator=g)
    lens = torch.tensor([T, T - 1, T // 2 + 3, 17])
    y = F.conv1d(x, w1, b1, padding=dil, dilation=dil) + cond
    gate, filt = torch.chunk(y, 2, dim=1)
    mask = (torch.arange(T)[None, :] < lens[:, None]).float()[:, None, :]
    z_ref = torch.sigmoid(gate) * torch.tanh(filt) * mask
    perm = torch.empty(2 * C, dtype=torch.long)
    perm[0::2] = torch.arange(C)
    perm[1::2] = torch.arange(C) + C
    planes = ops.split_f16(_cl(x))
    out, zp = ops.conv1d_umma_cl(planes, ops.pack_conv_weight_split(w1, interleave_halves=True, device="cuda"), 2 * C,
                                 bias=b1[perm].cuda(), K=3, dil=dil, pad=dil, act=ops.ACT_GATE,
                                 addend=_cl(cond[:, perm]), out_len=lens.cuda(), emit_planes=True,
                                 plane_add=padd.cuda())
    assert float((_bct(out) - z_ref).abs().max()) < 2e-5
    z = zp[0].float() + zp[1].float()
    assert float((_bct(z) - (z_ref + padd[None, :, None])).abs().max()) < 2e-5


@pytest.mark.parametrize("B,T", [(2, 333), (4, 2600)])
def test_conv1d_umma_dual_residual_from_planes(ops, monkeypatch, B, T):
    """DiffNet output projection with the residual stream kept ONLY as operand planes: h = hi + lo - step_emb[l] is
    recovered in the epilogue, (h + W_r z)/sqrt(2) + step_emb[l+1] is written back over the same planes (no fp32 h),
    the skip half accumulates as before.  Always takes the CTA-pair kernel."""
    _select(monkeypatch, "stream_tma")  # even with the pair kernel disabled by env this launch must take it
    g = torch.Generator().manual_seed(21)
    C = 256
    z = torch.rand(B, C, T, generator=g) * 2 - 1
    h = torch.randn(B, C, T, generator=g)
    emb, nxt = torch.randn(C, generator=g), torch.randn(C, generator=g)
    skip_old = torch.randn(B, C, T, generator=g)
    w2 = torch.randn(2 * C, C, 1, generator=g) / math.sqrt(C)
    b2 = torch.randn(2 * C, generator=g)
    o = F.conv1d(z, w2, b2)
    h_ref = (h + o[:, :C]) / math.sqrt(2.0)
    skip_ref = skip_old + o[:, C:]
    zp = ops.split_f16(_cl(z))
    yp = ops.split_f16(_cl(h), emb.cuda())  # y = h + step_emb[l]
    sbuf = _cl(skip_old)
    (o1, p1), _ = ops.conv1d_umma_dual_cl(
        zp, ops.pack_conv_weight_split(w2, device="cuda"), C,
        dict(bias=b2[:C].cuda(), res_planes=yp, res_plane_sub=emb.cuda(), out_div=math.sqrt(2.0), write_f32=False,
             emit_planes=True, out_planes=yp, plane_add=nxt.cuda()),
        dict(bias=b2[C:].cuda(), out=sbuf, beta=1.0))
    assert o1 is None and p1[0].data_ptr() == yp[0].data_ptr()
    y_next = yp[0].float() + yp[1].float()
    assert float((_bct(y_next) - (h_ref + nxt[None, :, None])).abs().max()) < 3e-5
    assert float((_bct(sbuf) - skip_ref).abs().max()) < 2e-5


@pytest.mark.parametrize("Cin,Cout", [(256, 1024), (1024, 256)])
def test_conv1d_umma_chunked_accumulation_is_fp32_class(ops, Cin, Cout):
    """conv1d impl 3 (the text encoder's k9 feed-forward convs, multi_layer_conv.py:52-67): chunks of 16 tensor-core
    accumulations summed in round-to-nearest fp32.  Against a float64 reference its error must be in the class of the
    fp32 CUDA-core kernel's (it feeds the integer duration rounding) and well below the plain tcgen05 accumulation's."""
    g = torch.Generator().manual_seed(Cin)
    B, T, K = 4, 256, 9
    lens = torch.tensor([256, 131, 200, 77])
    x = torch.randn(B, Cin, T, generator=g)
    x = x * (torch.arange(T)[None, None] < lens[:, None, None])
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = torch.relu(F.conv1d(x.double(), w.double(), b.double(), padding=4)) * (torch.arange(T)[None, None] < lens[:, None, None])
    planes = ops.split_f16(_cl(x))
    ws = ops.pack_conv_weight_split(w, device="cuda")
    kw = dict(bias=b.cuda(), K=K, dil=1, pad=4, act=ops.ACT_RELU, out_len=lens.cuda())
    out3, pl3 = ops.conv1d_umma_cl(planes, ws, Cout, impl=3, emit_planes=True, **kw)
    out2, _ = ops.conv1d_umma_cl(planes, ws, Cout, impl=2, **kw)
    out1 = ops.conv1d_cl(_cl(x), ops.pack_conv_weight(w, device="cuda"), Cout, impl=1, **kw)
    e3 = float((_bct(out3).double() - ref).abs().max())
    e2 = float((_bct(out2).double() - ref).abs().max())
    e1 = float((_bct(out1).double() - ref).abs().max())
    m3 = float((_bct(out3).double() - ref).abs().mean())
    m1 = float((_bct(out1).double() - ref).abs().mean())
    print(f"k9 {Cin}->{Cout}: max-abs err vs float64: fp32 CUDA cores {e1:.2e} (mean {m1:.2e}), tcgen05 chunked {e3:.2e} "
          f"(mean {m3:.2e}), tcgen05 plain {e2:.2e}")
    assert e3 < 3.0 * e1 + 2e-7 and m3 < 3.0 * m1 + 5e-8
    rec = pl3[0].float() + pl3[1].float()
    assert float((rec - out3).abs().max()) < 1e-6          # the emitted operand planes carry the same values
    assert float(_bct(out3)[1, :, 131:].abs().max()) == 0  # rows past the utterance are zero (out_len mask)


AA_CONV_CASES = [
    # C, K, dil, B, T: the BigVGAN geometries of the 32 / 64-channel stages; T not a multiple of the 128-row tile,
    # utterances shorter than one tile, halos up to 50 rows crossing both utterance ends
    (32, 3, 1, 2, 1000), (32, 3, 5, 2, 3001), (32, 7, 3, 1, 5000), (32, 11, 1, 3, 2999), (32, 11, 5, 2, 40000),
    (32, 11, 5, 3, 77), (32, 7, 1, 2, 5),
    (64, 3, 1, 2, 1000), (64, 3, 5, 2, 3001), (64, 7, 5, 2, 20000), (64, 7, 3, 3, 2999), (64, 7, 1, 2, 131),
]


@pytest.mark.parametrize("C,K,dil,B,T", AA_CONV_CASES)
def test_aa_conv_fused_is_bit_identical_to_two_launches(ops, C, K, dil, B, T):
    """AA-Snake fused into the conv's producer warps (pttspp_aa_conv1d_cl) == activation launch + conv launch, bit for
    bit, with residual / running-sum / division epilogue as BigVGAN's conv2 uses them; and against fp32 torch."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle import oracle
    from promptttspp_b200.layers.activations import AntiAliasActivation

    g = torch.Generator().manual_seed(C * 7 + K + dil)
    x = torch.randn(B, C, T, generator=g) * 2
    w = torch.randn(C, C, K, generator=g) / math.sqrt(C * K)
    b = torch.randn(C, generator=g)
    res = torch.randn(B, C, T, generator=g)
    prev = torch.randn(B, C, T, generator=g)
    act = AntiAliasActivation(C)
    alpha = torch.rand(C, generator=g) - 0.5
    up, down = act.up.filter.view(-1).cuda(), act.down.lowpass.filter.view(-1).cuda()
    pad = (K * dil - dil) // 2
    wsp = ops.pack_conv_weight_split(w, device="cuda")
    xc = _cl(x)
    # two launches: activation -> operand planes -> conv
    y = ops.aa_snake_cl(xc, alpha.cuda(), up, down, pair=True)
    out2, _ = ops.conv1d_umma_cl(ops.split_f16(y), wsp, C, bias=b.cuda(), K=K, dil=dil, pad=pad, res=_cl(res), beta=1.0,
                                 out=_cl(prev).clone(), out_div=3.0)
    out1 = ops.aa_conv1d_cl(xc, alpha.cuda(), up, down, wsp, bias=b.cuda(), K=K, dil=dil, pad=pad, res=_cl(res), beta=1.0,
                            out=_cl(prev).clone(), out_div=3.0)
    torch.cuda.synchronize()
    assert torch.equal(out1, out2), float((out1 - out2).abs().max())
    a = oracle.aa_activation(x, alpha.view(1, -1, 1), act.up.filter, act.down.lowpass.filter)
    ref = (res + F.conv1d(a, w, b, padding=pad, dilation=dil) + prev) / 3.0
    err = float((_bct(out1) - ref).abs().max())
    print(f"aa+conv {C} k{K} d{dil} T{T}: max-abs err vs fp32 torch {err:.3e}")
    assert err < 2e-5 + 6e-9 * K * C, err


@pytest.mark.parametrize("Cin,Cout,K,dil,B,T", [(128, 128, 11, 5, 2, 20000), (128, 128, 7, 1, 2, 20000),
                                                (256, 256, 11, 3, 2, 10000), (256, 256, 7, 5, 2, 10000)])
def test_conv1d_umma_impl4_pair_equals_streaming(ops, monkeypatch, Cin, Cout, K, dil, B, T):
    """impl 4 (the vocoder's convs): long contractions take the CTA-pair kernel when the batch fills the machine and the
    un-chunked streaming kernel otherwise -- same bits either way (an utterance synthesised alone equals the same
    utterance inside a batch), fp32 class against torch."""
    from promptttspp_b200 import _abi

    g = torch.Generator().manual_seed(Cin + K + dil)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = torch.randn(Cout, generator=g)
    pad = (K * dil - dil) // 2
    planes = ops.split_f16(_cl(x))
    wsp = ops.pack_conv_weight_split(w, device="cuda")
    n0 = _abi.lib().pttspp_launch_count()
    out_pair, _ = ops.conv1d_umma_cl(planes, wsp, Cout, bias=b.cuda(), K=K, dil=dil, pad=pad, impl=4)
    monkeypatch.setenv("PTTSPP_UMMA_PAIR", "0")
    _abi.lib().pttspp_debug_reload_env()
    out_stream, _ = ops.conv1d_umma_cl(planes, wsp, Cout, bias=b.cuda(), K=K, dil=dil, pad=pad, impl=4)
    torch.cuda.synchronize()
    assert _abi.lib().pttspp_launch_count() - n0 == 2
    assert torch.equal(out_pair, out_stream), float((out_pair - out_stream).abs().max())
    ref = F.conv1d(x, w, b, padding=pad, dilation=dil)
    err = float((_bct(out_pair) - ref).abs().max())
    print(f"impl4 {Cin}->{Cout} k{K} d{dil}: max-abs err {err:.3e}")
    assert err < 1e-5 + 1.2e-8 * K * Cin, err


@pytest.mark.parametrize("C,K,dil,B,T", [(32, 11, 5, 2, 40000), (32, 3, 1, 3, 13000), (32, 7, 3, 2, 20001),
                                         (64, 7, 5, 2, 20000), (64, 3, 1, 3, 13000), (64, 11, 5, 4, 12100),
                                         (64, 11, 1, 3, 13001)])
def test_conv1d_umma_wres_pair_equals_single_cta(ops, C, K, dil, B, T):
    """impl 4 with >= 2 blocks per SM: the CTA-pair form of the weight-resident kernel (cta_group::2, each CTA holds half
    of the stacked weight operand) == the single-CTA kernels (impl 2) bit for bit, incl. residual / running sum / division,
    an odd number of 128-row blocks (the last unit is half empty) and the 64-channel k = 11 case only the pair form can
    hold resident; fp32 class against torch."""
    from promptttspp_b200 import _abi

    g = torch.Generator().manual_seed(C + K + dil)
    x = torch.randn(B, C, T, generator=g)
    w = torch.randn(C, C, K, generator=g) / math.sqrt(C * K)
    b = torch.randn(C, generator=g)
    res = torch.randn(B, C, T, generator=g)
    prev = torch.randn(B, C, T, generator=g)
    pad = (K * dil - dil) // 2
    planes = ops.split_f16(_cl(x))
    wsp = ops.pack_conv_weight_split(w, device="cuda")
    kw = dict(bias=b.cuda(), K=K, dil=dil, pad=pad, res=_cl(res), beta=1.0, out_div=3.0)
    out4, _ = ops.conv1d_umma_cl(planes, wsp, C, out=_cl(prev).clone(), impl=4, **kw)
    out2, _ = ops.conv1d_umma_cl(planes, wsp, C, out=_cl(prev).clone(), impl=2, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out4, out2), float((out4 - out2).abs().max())
    ref = (res + F.conv1d(x, w, b, padding=pad, dilation=dil) + prev) / 3.0
    err = float((_bct(out4) - ref).abs().max())
    print(f"wres pair {C} k{K} d{dil}: max-abs err {err:.3e}")
    assert err < 1e-5 + 6e-9 * K * C, err
