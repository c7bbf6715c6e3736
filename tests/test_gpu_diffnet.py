"""GPU: the fused DiffNet residual-layer kernel (csrc/diffnet_layer.cu, through the C ABI) against a float64 torch
restatement of ResidualBlock.forward (promptttspp/modules/denoiser.py:69-83): gate output z (debug tap), residual planes,
skip sum; one layer per launch, the whole stack in one launch (flag-chained), ragged shapes."""
import math

import pytest
import torch
import torch.nn.functional as F

from promptttspp_b200 import ops

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
C_ = 256


def _make_layers(n_layers, seed, dils=None):
    g = torch.Generator().manual_seed(seed)
    layers = []
    for l in range(n_layers):
        layers.append(dict(
            dilated=torch.randn(2 * C_, C_, 3, generator=g) / math.sqrt(3 * C_),
            dilated_bias=torch.randn(2 * C_, generator=g) * 0.1,
            outp=torch.randn(2 * C_, C_, 1, generator=g) / math.sqrt(C_),
            outp_bias=torch.randn(2 * C_, generator=g) * 0.1,
            dil=(dils[l] if dils else 2 ** (l % 4)),
        ))
    return layers


def _reference(layers, h, cond, step, l0, l1, skip=None):
    """float64: returns (h', skip', z of the last layer); h [B, T, C], cond [L, B, T, 2C] (reference channel order)."""
    h = h.double()
    skip = None if skip is None else skip.double()
    z = None
    for l in range(l0, l1):
        L = layers[l]
        y = (h + step[l].double()).transpose(1, 2)
        pre = F.conv1d(y, L["dilated"].double().cuda(), L["dilated_bias"].double().cuda(), dilation=L["dil"],
                       padding=L["dil"]).transpose(1, 2) + cond[l].double()
        gate, filt = pre.chunk(2, dim=-1)
        z = torch.sigmoid(gate) * torch.tanh(filt)
        o = F.conv1d(z.transpose(1, 2), L["outp"].double().cuda(), L["outp_bias"].double().cuda()).transpose(1, 2)
        r, s = o.chunk(2, dim=-1)
        h = (h + r) / math.sqrt(2.0)
        skip = s if (skip is None or l == 0) else skip + s
    return h, skip, z


def _setup(B, T, n_layers, seed, dils=None):
    layers = _make_layers(n_layers, seed, dils)
    g = torch.Generator().manual_seed(seed + 1)
    h = torch.randn(B, T, C_, generator=g).cuda()
    cond = (torch.randn(n_layers, B, T, 2 * C_, generator=g) * 0.5).cuda()
    step = (torch.randn(n_layers + 1, C_, generator=g) * 0.3).cuda()
    stack = ops.DiffNetStack(layers, "cuda")
    return layers, stack, h, cond, step


def _planes(B, T):
    return tuple(torch.zeros(B, T, C_, dtype=torch.float16, device="cuda") for _ in range(2))


@pytest.mark.parametrize("B,T,dil", [(2, 256, 1), (3, 300, 2), (1, 77, 8), (5, 128, 4)])
def test_single_layer_matches_float64(B, T, dil):
    layers, stack, h, cond, step = _setup(B, T, 2, seed=B * 100 + T, dils=[dil, dil])
    cond_i = ops.DiffNetStack.interleave_cond(cond)
    y0 = ops.split_f16(h, step[0].contiguous())
    y1 = _planes(B, T)
    skip = torch.full((B, T, C_), float("nan"), device="cuda")  # layer 0 must overwrite, not accumulate
    z = torch.zeros(B, T, C_, device="cuda")
    done = stack.new_flags(B, T)
    stack.run(cond_i, step, (y0, y1), skip, done, epoch=1, layer_begin=0, layer_end=1, dbg_z=z)
    torch.cuda.synchronize()
    h_ref, s_ref, z_ref = _reference(layers, h, cond, step, 0, 1)
    ez = float((z.double() - z_ref).abs().max())
    hn = y1[0].double() + y1[1].double() - step[1].double()
    eh = float((hn - h_ref).abs().max())
    es = float((skip.double() - s_ref).abs().max())
    print(f"B={B} T={T} dil={dil}: z err {ez:.2e}, h' err {eh:.2e}, skip err {es:.2e}")
    assert ez < 2e-5 and eh < 2e-5 and es < 2e-5


def test_stack_in_one_launch_equals_layer_by_layer_and_float64():
    """20 layers, dilation cycle 1,2,4,8: the flag-chained single launch must be bit-identical to 20 single-layer
    launches (same arithmetic, different scheduling) and match float64; the last layer emits the skip planes."""
    B, T, NL = 4, 700, 20
    layers, stack, h, cond, step = _setup(B, T, NL, seed=5)
    cond_i = ops.DiffNetStack.interleave_cond(cond)
    outs = []
    for mode in ("chained", "per_layer"):
        y0 = tuple(t.clone() for t in ops.split_f16(h, step[0].contiguous()))
        y1 = _planes(B, T)
        skip = torch.zeros(B, T, C_, device="cuda")
        sp = _planes(B, T)
        done = stack.new_flags(B, T)
        if mode == "chained":
            stack.run(cond_i, step, (y0, y1), skip, done, epoch=1, skip_planes=sp)
        else:
            for l in range(NL):
                stack.run(cond_i, step, (y0, y1), skip, done, epoch=l + 1, layer_begin=l, layer_end=l + 1, skip_planes=sp)
        torch.cuda.synchronize()
        outs.append((y0, y1, skip.clone(), sp))
    a, b = outs
    assert torch.equal(a[3][0], b[3][0]) and torch.equal(a[3][1], b[3][1])  # skip planes
    assert torch.equal(a[0][0], b[0][0]) and torch.equal(a[1][0], b[1][0])
    h_ref, s_ref, _ = _reference(layers, h, cond, step, 0, NL)
    s = a[3][0].double() + a[3][1].double()
    es = float((s - s_ref).abs().max())
    # the residual stream after layer NL-2 (the last layer has no residual output): planes y[(NL-1) & 1]
    h_ref18, _, _ = _reference(layers, h, cond, step, 0, NL - 1)
    yl = a[(NL - 1) & 1]
    eh = float((yl[0].double() + yl[1].double() - step[NL - 1].double() - h_ref18).abs().max())
    print(f"20-layer stack: skip-sum err {es:.2e} (|skip| max {float(s_ref.abs().max()):.2f}), h err {eh:.2e}")
    assert es < 2e-4 and eh < 5e-5


def test_stack_repeated_epochs_full_machine():
    """cfg2-sized rows (16 x 2048): every cluster busy, several epochs over the same flag array (as the sampling loop
    does); results of all epochs identical."""
    B, T, NL = 16, 2048, 8
    layers, stack, h, cond, step = _setup(B, T, NL, seed=9)
    cond_i = ops.DiffNetStack.interleave_cond(cond)
    done = stack.new_flags(B, T)
    ref = None
    for epoch in range(1, 4):
        y0 = tuple(t.clone() for t in ops.split_f16(h, step[0].contiguous()))
        y1 = _planes(B, T)
        skip = torch.zeros(B, T, C_, device="cuda")
        sp = _planes(B, T)
        stack.run(cond_i, step, (y0, y1), skip, done, epoch=epoch, skip_planes=sp)
        torch.cuda.synchronize()
        cur = (sp[0].clone(), sp[1].clone(), y0[0].clone(), y1[0].clone())
        if ref is None:
            ref = cur
        else:
            assert all(torch.equal(x, y) for x, y in zip(ref, cur))
    h_ref, s_ref, _ = _reference(layers, h[:2], cond[:, :2], step, 0, NL)
    es = float((ref[0][:2].double() + ref[1][:2].double() - s_ref).abs().max())
    print(f"8-layer stack at 32768 rows: skip-sum err {es:.2e}")
    assert es < 1e-4
