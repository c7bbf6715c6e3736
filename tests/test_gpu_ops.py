"""GPU: every op-level C-ABI entry point against plain torch fp32 on the CPU (the oracle's primitives)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import oracle

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def ops():
    from promptttspp_b200 import _abi, ops

    _abi.check(_abi.lib().pttspp_device_check())
    return ops


def _cl(x):  # [B, C, T] -> channels-last cuda
    return x.transpose(1, 2).contiguous().cuda()


def _bct(x):  # channels-last cuda -> [B, C, T] cpu
    return x.cpu().transpose(1, 2)


def _mask(lens, T):
    return (torch.arange(T)[None] < lens[:, None]).float().unsqueeze(1)


CONV_CASES = [
    # Cin, Cout, K, dil, B, T
    (256, 512, 3, 4, 2, 300),
    (80, 256, 1, 1, 3, 77),
    (256, 80, 1, 1, 2, 130),
    (32, 1, 7, 1, 2, 1000),
    (1024, 256, 9, 1, 2, 50),
    (64, 64, 11, 5, 1, 515),
    (32, 32, 3, 3, 2, 700),
    (128, 128, 7, 1, 1, 129),
    (16, 2, 1, 1, 1, 5),
]


@pytest.mark.parametrize("Cin,Cout,K,dil,B,T", CONV_CASES)
def test_conv1d_plain(ops, Cin, Cout, K, dil, B, T):
    g = torch.Generator().manual_seed(Cin * 7 + Cout + K)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = torch.randn(Cout, generator=g)
    pad = (K * dil - dil) // 2
    ref = F.conv1d(x, w, b, padding=pad, dilation=dil)
    out = ops.conv1d_cl(_cl(x), ops.pack_conv_weight(w, device="cuda"), Cout, bias=b.cuda(), K=K, dil=dil, pad=pad,
                        impl=1)
    assert torch.allclose(_bct(out), ref, atol=2e-5, rtol=1e-5), float((_bct(out) - ref).abs().max())


def test_conv1d_fused_epilogues(ops):
    g = torch.Generator().manual_seed(5)
    B, C, T = 3, 256, 200
    lens = torch.tensor([200, 131, 7])
    m = _mask(lens, T)
    x = torch.randn(B, C, T, generator=g)
    res = torch.randn(B, C, T, generator=g)
    w = torch.randn(C, C, 9, generator=g) / math.sqrt(C * 9)
    b = torch.randn(C, generator=g)
    # FFN second half: res + 0.5 * conv(x * m) * m  (in_len, out_len, residual, alpha)
    ref = res + 0.5 * F.conv1d(x * m, w, b, padding=4) * m
    out = ops.conv1d_cl(_cl(x), ops.pack_conv_weight(w, device="cuda"), C, bias=b.cuda(), K=9, pad=4,
                        in_len=lens.cuda(), out_len=lens.cuda(), res=_cl(res), alpha=0.5, impl=1)
    assert torch.allclose(_bct(out), ref, atol=2e-5)
    # activations
    for act, fn in ((ops.ACT_RELU, torch.relu), (ops.ACT_GELU, F.gelu), (ops.ACT_SWISH, lambda t: t * torch.sigmoid(t)),
                    (ops.ACT_TANH, torch.tanh)):
        out = ops.conv1d_cl(_cl(x), ops.pack_conv_weight(w, device="cuda"), C, bias=b.cuda(), K=9, pad=4, act=act,
                            impl=1)
        assert torch.allclose(_bct(out), fn(F.conv1d(x, w, b, padding=4)), atol=2e-5), act
    # accumulate + divide: (res + conv + old) / 3
    old = torch.randn(B, C, T, generator=g)
    outbuf = _cl(old)
    ops.conv1d_cl(_cl(x), ops.pack_conv_weight(w, device="cuda"), C, bias=b.cuda(), K=9, pad=4, res=_cl(res), beta=1.0,
                  out_div=3.0, out=outbuf, impl=1)
    assert torch.allclose(_bct(outbuf), (res + F.conv1d(x, w, b, padding=4) + old) / 3, atol=2e-5)
    # in-place residual (out aliases res), scaled accumulator
    buf = _cl(res)
    ops.conv1d_cl(_cl(x), ops.pack_conv_weight(w, device="cuda"), C, bias=b.cuda(), K=9, pad=4, res=buf, out=buf,
                  out_div=math.sqrt(2.0), acc_scale=0.25, impl=1)
    assert torch.allclose(_bct(buf), (res + 0.25 * F.conv1d(x, w, None, padding=4) + b[None, :, None]) / math.sqrt(2.0),
                          atol=2e-5)


def test_conv1d_diffnet_gate(ops):
    """dilated conv of (x + step) with zero padding + conditioner addend -> sigmoid(gate) * tanh(filter)."""
    g = torch.Generator().manual_seed(6)
    B, C, T, dil = 2, 256, 150, 8
    x = torch.randn(B, C, T, generator=g)
    step = torch.randn(C, generator=g)
    cond = torch.randn(B, 2 * C, T, generator=g)
    w = torch.randn(2 * C, C, 3, generator=g) / math.sqrt(C * 3)
    b = torch.randn(2 * C, generator=g)
    y = F.conv1d(x + step[None, :, None], w, b, padding=dil, dilation=dil) + cond
    gate, filt = torch.chunk(y, 2, dim=1)
    ref = torch.sigmoid(gate) * torch.tanh(filt)
    perm = torch.empty(2 * C, dtype=torch.long)
    perm[0::2] = torch.arange(C)
    perm[1::2] = torch.arange(C) + C
    out = ops.conv1d_cl(_cl(x), ops.pack_conv_weight(w, interleave_halves=True, device="cuda"), 2 * C,
                        bias=b[perm].cuda(), K=3, dil=dil, pad=dil, act=ops.ACT_GATE, in_add=step.cuda(),
                        addend=_cl(cond[:, perm]), impl=1)
    assert out.shape == (B, T, C)
    assert torch.allclose(_bct(out), ref, atol=2e-5)


@pytest.mark.parametrize("Cin,Cout,k,s", [(512, 256, 12, 6), (256, 128, 10, 5), (128, 64, 8, 4), (64, 32, 4, 2)])
def test_polyphase_conv_transpose(ops, Cin, Cout, k, s):
    g = torch.Generator().manual_seed(k)
    B, T = 2, 37
    x = torch.randn(B, Cin, T, generator=g)
    v = torch.randn(Cin, Cout, k, generator=g) / math.sqrt(Cin * 2)
    gg = v.flatten(1).norm(dim=1).view(Cin, 1, 1) * (torch.rand(Cin, 1, 1, generator=g) + 0.5)
    b = torch.randn(Cout, generator=g)
    p, op = s // 2 + s % 2, s % 2
    ref = F.conv_transpose1d(x, torch._weight_norm(v, gg, 0), b, stride=s, padding=p, output_padding=op)
    Lout = ref.shape[-1]
    assert Lout == T * s
    packed = ops.pack_convtr_weight(v, s, gg, device="cuda")
    J = k // s
    out = torch.full((B, Lout, Cout), float("nan"), device="cuda")
    xin = _cl(x)
    for r in range(s):
        off = r - p
        m_begin = -(off // s) if off < 0 else 0  # ceil(-off / s)
        m_end = (Lout - 1 - off) // s
        ops.conv1d_cl(xin, packed[r], Cout, bias=b.cuda(), K=J, dil=1, pad=J - 1, T_out=Lout, m_begin=m_begin,
                      M=m_end - m_begin + 1, out_mul=s, out_off=off, out=out, impl=1)
    assert not torch.isnan(out).any(), "every output row must be written by exactly one phase"
    assert torch.allclose(_bct(out), ref, atol=2e-5)


def test_layernorm_variants(ops):
    g = torch.Generator().manual_seed(8)
    B, T, C = 3, 97, 256
    lens = torch.tensor([97, 50, 1])
    x = torch.randn(B, T, C, generator=g) * 3 + 1
    x2 = torch.randn(B, T, C, generator=g)
    pe = torch.randn(T, C, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    m = (torch.arange(T)[None] < lens[:, None]).float().unsqueeze(-1)
    out = ops.layernorm_cl(x.cuda(), gamma.cuda(), beta.cuda(), 1e-12)
    assert torch.allclose(out.cpu(), F.layer_norm(x, (C,), gamma, beta, 1e-12), atol=2e-5)
    out = ops.layernorm_cl(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, in2=x2.cuda(), out_len=lens.cuda())
    assert torch.allclose(out.cpu(), F.layer_norm(x + x2, (C,), gamma, beta, 1e-5) * m, atol=2e-5)
    out = ops.layernorm_cl(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, row_add=pe.cuda(), in_scale=16.0,
                           in_len=lens.cuda())
    assert torch.allclose(out.cpu(), F.layer_norm(x * m * 16.0 + pe, (C,), gamma, beta, 1e-5), atol=2e-5)
    # all-zero rows with eps 1e-12 must give beta, not NaN (padded phoneme rows of the Conformer)
    z = torch.zeros(1, 4, C)
    out = ops.layernorm_cl(z.cuda(), gamma.cuda(), beta.cuda(), 1e-12)
    assert torch.allclose(out.cpu(), beta.expand(1, 4, C))
    # in place
    buf = x.clone().cuda()
    from promptttspp_b200 import _abi
    import ctypes as C_
    d = _abi.LayerNormDesc()
    d.in_ = buf.data_ptr(); d.gamma = gamma.cuda().data_ptr(); d.beta = beta.cuda().data_ptr(); d.out = buf.data_ptr()
    gam, bet = gamma.cuda(), beta.cuda()
    d.gamma, d.beta = gam.data_ptr(), bet.data_ptr()
    d.bs = T * C; d.ld = C; d.B = B; d.T = T; d.C = C; d.eps = 1e-5; d.in_scale = 1.0
    _abi.check(_abi.lib().pttspp_layernorm_cl(C_.byref(d), _abi.stream_ptr()))
    assert torch.allclose(buf.cpu(), F.layer_norm(x, (C,), gamma, beta, 1e-5), atol=2e-5)


@pytest.mark.parametrize("B,C,L", [(2, 32, 1), (1, 64, 3), (2, 32, 7), (1, 96, 37), (2, 256, 64), (1, 32, 1001),
                                    (1, 512, 130)])
def test_aa_snake(ops, B, C, L):
    from promptttspp_b200.layers.activations import AntiAliasActivation

    g = torch.Generator().manual_seed(L)
    act = AntiAliasActivation(C)
    x = torch.randn(B, C, L, generator=g) * 2
    alpha = torch.rand(1, C, 1, generator=g) - 0.5
    ref = oracle.aa_activation(x, alpha, act.up.filter, act.down.lowpass.filter)
    out = ops.aa_snake_cl(_cl(x), alpha.view(-1).cuda(), act.up.filter.view(-1).cuda(),
                          act.down.lowpass.filter.view(-1).cuda())
    assert torch.allclose(_bct(out), ref, atol=2e-5), float((_bct(out) - ref).abs().max())
    # the channel-pair kernel the BigVGAN handle runs (packed fp32x2, rolling strips): bit-identical per channel
    pair = ops.aa_snake_cl(_cl(x), alpha.view(-1).cuda(), act.up.filter.view(-1).cuda(),
                           act.down.lowpass.filter.view(-1).cuda(), pair=True)
    assert torch.equal(pair, out), float((pair - out).abs().max())


@pytest.mark.parametrize("B,C,L", [(2, 64, 63), (1, 32, 64), (3, 32, 65), (1, 128, 129), (2, 32, 4097)])
def test_aa_snake_pair_kernel_strip_boundaries(ops, B, C, L):
    """strip (64 outputs) and block (8 outputs) boundaries of the rolling channel-pair kernel against the oracle"""
    from promptttspp_b200.layers.activations import AntiAliasActivation

    g = torch.Generator().manual_seed(L + C)
    act = AntiAliasActivation(C)
    x = torch.randn(B, C, L, generator=g) * 2
    alpha = torch.rand(1, C, 1, generator=g) - 0.5
    ref = oracle.aa_activation(x, alpha, act.up.filter, act.down.lowpass.filter)
    out = ops.aa_snake_cl(_cl(x), alpha.view(-1).cuda(), act.up.filter.view(-1).cuda(),
                          act.down.lowpass.filter.view(-1).cuda(), pair=True)
    assert torch.allclose(_bct(out), ref, atol=2e-5), float((_bct(out) - ref).abs().max())


def test_duration_quantize_and_length_regulator_bit_exact(ops):
    g = torch.Generator().manual_seed(11)
    B, Tx, C = 4, 57, 256
    lens = torch.tensor([57, 30, 1, 44])
    log_d = torch.randn(B, Tx, generator=g) * 0.8 + 1.5
    # values next to (not on) the .5 ties: an exact tie is decided by the last ulp of exp(), which differs between
    # any two libm implementations; the kernel uses the correctly rounded exp (double -> float)
    log_d[0, :8] = torch.tensor([0.49, 0.51, 1.499, 1.501, 2.499, 2.501, 0.05, 1.0]).log()
    log_d[1, :2] = torch.tensor([-3.0, 12.0])  # clamp_min(1); a very long phoneme
    pm = (torch.arange(Tx)[None] < lens[:, None])
    ref_d, ref_len = oracle.quantize_durations(log_d.unsqueeze(1), pm.unsqueeze(1).long())
    dur, flen = ops.duration_quantize(log_d.cuda(), lens.cuda())
    assert torch.equal(dur.cpu(), ref_d.squeeze(1)) and torch.equal(flen.cpu(), ref_len)
    Ty = int(ref_len.max())
    x = torch.randn(B, C, Tx, generator=g)
    fm = (torch.arange(Ty)[None] < ref_len[:, None]).float().unsqueeze(1)
    ref = oracle.length_regulate_dense(x, ref_d.float(), pm.unsqueeze(1).float(), fm)
    out, idx = ops.length_regulate(_cl(x), dur, Ty)
    assert torch.equal(_bct(out), ref), "the gather must reproduce the one-hot matmul exactly"
    assert torch.equal(idx.cpu().long(), oracle.length_regulate_indices(ref_d.squeeze(1), Ty))
    # longer than one scan chunk per thread, with a wider output than needed (trailing zeros)
    Tx2 = 700
    d2 = torch.randint(0, 5, (2, Tx2), generator=g)
    x2 = torch.randn(2, 32, Tx2, generator=g)
    Ty2 = int(d2.sum(1).max()) + 9
    out2, idx2 = ops.length_regulate(_cl(x2), d2.cuda(), Ty2)
    assert torch.equal(idx2.cpu().long(), oracle.length_regulate_indices(d2, Ty2))
    sel = idx2.cpu().long().clamp(min=0)
    ref2 = torch.gather(x2, 2, sel.unsqueeze(1).expand(-1, 32, -1)) * (idx2.cpu() >= 0).unsqueeze(1)
    assert torch.equal(_bct(out2), ref2)


@pytest.mark.parametrize("legacy", [True, False])
@pytest.mark.parametrize("T,lens", [(64, [64, 33, 5]), (37, [37, 36, 1]), (256, [256, 129, 17]), (200, [131, 200, 128, 129]),
                                    (129, [129, 2]), (300, [300, 77])])
def test_relpos_attention(ops, legacy, T, lens):
    """T <= 256: the fused tcgen05 kernel (csrc/attention_umma.cu); T = 300: the CUDA-core fallback (csrc/attention.cu)."""
    g = torch.Generator().manual_seed(T + int(legacy))
    H, dk = 2, 128
    B = len(lens)
    lens = torch.tensor(lens)
    q, k, v = (torch.randn(B, T, H * dk, generator=g) for _ in range(3))
    Tp = T if legacy else 2 * T - 1
    p = torch.randn(Tp, H * dk, generator=g)
    bu, bv = torch.randn(H, dk, generator=g) * 0.3, torch.randn(H, dk, generator=g) * 0.3
    qh = q.view(B, T, H, dk)
    kh, vh = k.view(B, T, H, dk).transpose(1, 2), v.view(B, T, H, dk).transpose(1, 2)
    ph = p.view(1, Tp, H, dk).transpose(1, 2)
    ac = torch.matmul((qh + bu).transpose(1, 2), kh.transpose(-2, -1))
    bd = torch.matmul((qh + bv).transpose(1, 2), ph.transpose(-2, -1))
    bd = oracle.rel_shift_legacy(bd) if legacy else oracle.rel_shift_new(bd)
    pad = torch.arange(T)[None] < lens[:, None]
    m = ~(pad.unsqueeze(-2) & pad.unsqueeze(-1)).unsqueeze(1)
    scores = ((ac + bd) / math.sqrt(dk)).masked_fill(m, torch.finfo(torch.float32).min)
    attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
    ref = torch.matmul(attn, vh).transpose(1, 2).reshape(B, T, H * dk)
    out = ops.relpos_attention(q.cuda(), k.cuda(), v.cuda(), p.cuda(), bu.cuda(), bv.cuda(), lens.cuda(), H, legacy)
    err = float((out.cpu() - ref).abs().max())
    print(f"relpos attention T={T} legacy={legacy}: max-abs err {err:.2e}")
    assert torch.isfinite(out).all() and err < 2e-5, err


def test_lowpass_filter_matches_reference_golden(golden_dir):
    """pttspp_iir_filtfilt behind utils.model.lowpass_filter vs the reference function's output (app.py:77 path)."""
    import numpy as np
    from golden_cases import lowpass_inputs
    from promptttspp_b200.utils.model import lowpass_filter

    gold = np.load(golden_dir / "lowpass.npz")
    for i, x in enumerate(lowpass_inputs()):
        y = lowpass_filter(x.cuda(), 100, cutoff=20).cpu()
        ref = torch.from_numpy(gold[f"y{i}"])
        assert y.shape == ref.shape
        err = float((y - ref).abs().max())
        print(f"lowpass case {i}: max-abs err {err:.2e}")
        assert err < 2e-5
    with pytest.raises(RuntimeError):
        lowpass_filter(torch.zeros(1, 1, 100))  # CPU tensor: no host fallback


def test_lowpass_filter_ragged_rows_equal_per_utterance_calls():
    """a zero-padded batch filtered with `lengths` must give every row what a per-utterance call
    (app.py:76-77) gives on its own frames, incl. the short-input pass-through; the padding is left untouched."""
    from promptttspp_b200.utils.model import lowpass_filter

    g = torch.Generator().manual_seed(5)
    lens = [400, 131, 18, 19, 1, 257]
    B, T = len(lens), max(lens)
    x = torch.zeros(B, 1, T)
    for b, n in enumerate(lens):
        x[b, 0, :n] = 5.0 + 0.3 * torch.randn(n, generator=g)  # log-f0 like: a step of ~5 at the end of the row
    y = lowpass_filter(x.cuda(), 100, cutoff=20, lengths=torch.tensor(lens).cuda()).cpu()
    for b, n in enumerate(lens):
        single = lowpass_filter(x[b:b + 1, :, :n].contiguous().cuda(), 100, cutoff=20).cpu()
        assert torch.equal(y[b:b + 1, :, :n], single), (b, n, float((y[b:b + 1, :, :n] - single).abs().max()))
        assert torch.equal(y[b, :, n:], x[b, :, n:])
    # without lengths the tail of a short row is contaminated by the step into the padding (what the advice flagged)
    bad = lowpass_filter(x.cuda(), 100, cutoff=20).cpu()
    assert float((bad[1, 0, :131] - y[1, 0, :131]).abs().max()) > 0.5


@pytest.mark.parametrize("numel", [1, 255, 4096, 80 * 519, 16 * 80 * 2582, 5_000_001])
def test_philox_normal_is_torch_randn_bit_for_bit(ops, numel):
    """csrc/philox.cu draws the sampler's per-step noise natively; it must be THE torch draw (same values, same advance of
    the generator), so a seeded run is unchanged and matches a seeded GPU run of the reference (diffusion.py:218)."""
    torch.cuda.init()  # default_generators is populated lazily
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    for seed in (0, 1234567):
        torch.manual_seed(seed)
        torch.randn(1000, device="cuda")                      # move the offset off zero
        off0 = gen.get_offset()
        want = torch.randn(numel, device="cuda")
        off1 = gen.get_offset()
        got, adv = ops.philox_normal(numel, gen.initial_seed(), off0)
        assert adv == off1 - off0, (adv, off1 - off0)
        assert torch.equal(got, want), float((got - want).abs().max())
