"""Thin torch-tensor wrappers over the op-level C ABI (include/pttspp_b200.h).

Activations are channels-last float32 CUDA tensors x[b, t, c].  These wrappers allocate outputs with
torch and launch on torch's current stream; all arithmetic happens in libpttspp_b200.so.
"""
import ctypes as C

import torch

from . import _abi
from ._abi import ACT_GATE, ACT_GELU, ACT_NONE, ACT_RELU, ACT_SWISH, ACT_TANH  # noqa: F401


def _round4(n):
    return (n + 3) // 4 * 4


def pack_conv_weight(weight, g=None, interleave_halves=False, device=None):
    """torch Conv1d weight [Cout, Cin, K] (or Linear [Cout, Cin]) -> packed [K, Cin, w_ld] on `device`."""
    w = weight.detach().float().cpu().contiguous()
    if w.dim() == 2:
        w = w.unsqueeze(-1)
    Cout, Cin, K = w.shape
    w_ld = _round4(Cout)
    packed = torch.empty(K, Cin, w_ld, dtype=torch.float32)
    gp = None if g is None else g.detach().float().cpu().contiguous().view(-1)
    _abi.check(_abi.lib().pttspp_pack_conv_weight(_abi.ptr(w), _abi.ptr(gp), Cout, Cin, K, _abi.ptr(packed), w_ld,
                                                  int(interleave_halves), None))
    return packed.to(device) if device is not None else packed


def pack_convtr_weight(weight, stride, g=None, device=None):
    """torch ConvTranspose1d weight [Cin, Cout, Kt] -> polyphase [stride, Kt/stride, Cin, w_ld]."""
    w = weight.detach().float().cpu().contiguous()
    Cin, Cout, Kt = w.shape
    w_ld = _round4(Cout)
    packed = torch.empty(stride, Kt // stride, Cin, w_ld, dtype=torch.float32)
    gp = None if g is None else g.detach().float().cpu().contiguous().view(-1)
    _abi.check(_abi.lib().pttspp_pack_convtr_weight(_abi.ptr(w), _abi.ptr(gp), Cin, Cout, Kt, stride,
                                                    _abi.ptr(packed), w_ld, None))
    return packed.to(device) if device is not None else packed


def conv1d_cl(x, packed_w, Cout, bias=None, K=1, dil=1, pad=0, act=ACT_NONE, in_len=None, out_len=None,
              in_add=None, addend=None, res=None, res_scale=1.0, alpha=1.0, beta=0.0, out=None, acc_scale=1.0,
              out_div=0.0, T_out=None, m_begin=0, M=None, out_mul=1, out_off=0, in_stride=1, impl=0):
    _abi.require_cuda(x, "conv1d_cl")
    B, T, Cin = x.shape
    w_ld = packed_w.shape[-1]
    out_cols = Cout // 2 if act == ACT_GATE else Cout
    T_out = T if T_out is None else T_out
    if out is None:
        out = torch.zeros(B, T_out, out_cols, device=x.device, dtype=torch.float32)
    d = _abi.Conv1dDesc()
    d.in_ = x.data_ptr(); d.in_bs = x.stride(0); d.in_ld = x.stride(1); d.T_in = T; d.Cin = Cin
    d.w = packed_w.data_ptr(); d.w_ld = w_ld; d.bias = None if bias is None else bias.data_ptr()
    d.K = K; d.dil = dil; d.pad = pad; d.in_stride = in_stride
    d.out = out.data_ptr(); d.out_bs = out.stride(0); d.out_ld = out.stride(1); d.T_out = T_out; d.Cout = Cout
    d.m_begin = m_begin; d.M = T_out if M is None else M; d.out_mul = out_mul; d.out_off = out_off
    d.in_len = None if in_len is None else in_len.data_ptr()
    d.out_len = None if out_len is None else out_len.data_ptr()
    d.in_add = None if in_add is None else in_add.data_ptr()
    if addend is not None:
        d.addend = addend.data_ptr(); d.addend_bs = addend.stride(0); d.addend_ld = addend.stride(1)
    d.act = act; d.acc_scale = acc_scale
    if res is not None:
        d.res = res.data_ptr(); d.res_bs = res.stride(0); d.res_ld = res.stride(1)
    d.res_scale = res_scale; d.alpha = alpha; d.beta = beta; d.out_div = out_div
    d.B = B; d.impl = impl
    _abi.check(_abi.lib().pttspp_conv1d_cl(C.byref(d), _abi.stream_ptr(x.device)))
    return out


def layernorm_cl(x, gamma, beta, eps, in2=None, row_add=None, in_scale=1.0, in_len=None, out_len=None):
    _abi.require_cuda(x, "layernorm_cl")
    B, T, Cc = x.shape
    out = torch.empty_like(x)
    d = _abi.LayerNormDesc()
    d.in_ = x.data_ptr(); d.in2 = None if in2 is None else in2.data_ptr()
    d.row_add = None if row_add is None else row_add.data_ptr()
    d.gamma = gamma.data_ptr(); d.beta = beta.data_ptr(); d.out = out.data_ptr()
    d.bs = x.stride(0); d.ld = x.stride(1); d.B = B; d.T = T; d.C = Cc; d.eps = eps; d.in_scale = in_scale
    d.in_len = None if in_len is None else in_len.data_ptr()
    d.out_len = None if out_len is None else out_len.data_ptr()
    _abi.check(_abi.lib().pttspp_layernorm_cl(C.byref(d), _abi.stream_ptr(x.device)))
    return out


def aa_snake_cl(x, log_alpha, up_filter, down_filter, pair=False):
    """pair=True: the channel-pair kernel (symmetric filters, even C) the BigVGAN handle runs."""
    _abi.require_cuda(x, "aa_snake_cl")
    B, L, Cc = x.shape
    y = torch.empty_like(x)
    fn = _abi.lib().pttspp_aa_snake_pair_cl if pair else _abi.lib().pttspp_aa_snake_cl
    _abi.check(fn(_abi.ptr(x), _abi.ptr(y), B, L, Cc, _abi.ptr(log_alpha),
                                             _abi.ptr(up_filter), _abi.ptr(down_filter), _abi.stream_ptr(x.device)))
    return y


def duration_quantize(log_d, phone_len):
    _abi.require_cuda(log_d, "duration_quantize")
    B, Tx = log_d.shape
    dur = torch.empty(B, Tx, dtype=torch.int64, device=log_d.device)
    flen = torch.empty(B, dtype=torch.int64, device=log_d.device)
    _abi.check(_abi.lib().pttspp_duration_quantize(_abi.ptr(log_d), _abi.ptr(phone_len), B, Tx, _abi.ptr(dur),
                                                   _abi.ptr(flen), _abi.stream_ptr(log_d.device)))
    return dur, flen


def length_regulate(x, dur, Ty):
    _abi.require_cuda(x, "length_regulate")
    B, Tx, Cc = x.shape
    out = torch.empty(B, Ty, Cc, device=x.device, dtype=torch.float32)
    idx = torch.empty(B, Ty, dtype=torch.int32, device=x.device)
    _abi.check(_abi.lib().pttspp_length_regulate(_abi.ptr(x), _abi.ptr(dur), B, Tx, Cc, Ty, _abi.ptr(out),
                                                 _abi.ptr(idx), _abi.stream_ptr(x.device)))
    return out, idx


def relpos_attention(q, k, v, p, bias_u, bias_v, lens, heads, legacy):
    _abi.require_cuda(q, "relpos_attention")
    B, T, HD = q.shape
    dk = HD // heads
    Tp = p.shape[0]
    scratch = torch.empty(B * heads * T * Tp, device=q.device, dtype=torch.float32)
    out = torch.empty(B, T, HD, device=q.device, dtype=torch.float32)
    _abi.check(_abi.lib().pttspp_relpos_attention(
        _abi.ptr(q), _abi.ptr(k), _abi.ptr(v), _abi.ptr(p), _abi.ptr(bias_u), _abi.ptr(bias_v), _abi.ptr(lens), B, T,
        heads, dk, int(legacy), _abi.ptr(scratch), _abi.ptr(out), _abi.stream_ptr(q.device)))
    return out


# ---- split-fp16 operand planes (tcgen05 path) --------------------------------------------------

def split_f16(x, add=None):
    """fp32 CUDA tensor [..., C] -> (hi, lo) fp16 planes of (x + add[c])."""
    _abi.require_cuda(x, "split_f16")
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _abi.check(_abi.lib().pttspp_split_f16(_abi.ptr(x), _abi.ptr(add), x.numel(), x.shape[-1], _abi.ptr(hi),
                                           _abi.ptr(lo), _abi.stream_ptr(x.device)))
    return hi, lo


def pack_conv_weight_split(weight, g=None, interleave_halves=False, device=None):
    """torch Conv1d weight [Cout, Cin, K] -> (w_hi, w_lo [K, Cout, Cin] fp16, scale_inv)."""
    w = weight.detach().float().cpu().contiguous()
    if w.dim() == 2:
        w = w.unsqueeze(-1)
    Cout, Cin, K = w.shape
    hi = torch.empty(K, Cout, Cin, dtype=torch.float16)
    lo = torch.empty(K, Cout, Cin, dtype=torch.float16)
    gp = None if g is None else g.detach().float().cpu().contiguous().view(-1)
    sc = C.c_float(0.0)
    _abi.check(_abi.lib().pttspp_pack_conv_weight_split(_abi.ptr(w), _abi.ptr(gp), Cout, Cin, K, _abi.ptr(hi),
                                                        _abi.ptr(lo), int(interleave_halves), C.byref(sc)))
    if device is not None:
        hi, lo = hi.to(device), lo.to(device)
    return hi, lo, sc.value


def conv1d_umma_cl(x_planes, w_split, Cout, bias=None, K=1, dil=1, pad=0, act=ACT_NONE, out_len=None, addend=None,
                   res=None, res_scale=1.0, alpha=1.0, beta=0.0, out=None, acc_scale=1.0, out_div=0.0,
                   emit_planes=False, plane_add=None, write_f32=True, _desc_only=False, res_planes=None,
                   res_plane_sub=None, out_planes=None, impl=2):
    """tcgen05 path on pre-split operands.  x_planes = (hi, lo) [B, T, Cin] fp16; w_split from
    pack_conv_weight_split.  Returns out (fp32) and/or (out_hi, out_lo)."""
    xh, xl = x_planes
    wh, wl, scale_inv = w_split
    _abi.require_cuda(xh, "conv1d_umma_cl")
    B, T, Cin = xh.shape
    out_cols = Cout // 2 if act == ACT_GATE else Cout
    if out is None and write_f32:
        out = torch.zeros(B, T, out_cols, device=xh.device, dtype=torch.float32)
    d = _abi.Conv1dDesc()
    d.in_bs = xh.stride(0); d.in_ld = xh.stride(1); d.T_in = T; d.Cin = Cin
    d.in_hi = xh.data_ptr(); d.in_lo = xl.data_ptr(); d.w_hi = wh.data_ptr(); d.w_lo = wl.data_ptr()
    d.w_scale_inv = scale_inv
    d.bias = None if bias is None else bias.data_ptr()
    d.K = K; d.dil = dil; d.pad = pad; d.in_stride = 1
    if out is not None:
        d.out = out.data_ptr(); d.out_bs = out.stride(0); d.out_ld = out.stride(1)
    d.T_out = T; d.Cout = Cout; d.m_begin = 0; d.M = T; d.out_mul = 1; d.out_off = 0
    d.out_len = None if out_len is None else out_len.data_ptr()
    if addend is not None:
        d.addend = addend.data_ptr(); d.addend_bs = addend.stride(0); d.addend_ld = addend.stride(1)
    d.act = act; d.acc_scale = acc_scale
    if res is not None:
        d.res = res.data_ptr(); d.res_bs = res.stride(0); d.res_ld = res.stride(1)
    if res_planes is not None:  # residual = hi + lo - sub, read from operand planes (CTA-pair kernel only)
        d.res_hi = res_planes[0].data_ptr(); d.res_lo = res_planes[1].data_ptr()
        d.res_plane_bs = res_planes[0].stride(0); d.res_plane_ld = res_planes[0].stride(1)
        d.res_plane_sub = None if res_plane_sub is None else res_plane_sub.data_ptr()
    d.res_scale = res_scale; d.alpha = alpha; d.beta = beta; d.out_div = out_div
    d.B = B; d.impl = impl  # 2 = tcgen05, 3 = tcgen05 with chunked near-fp32 accumulation
    planes = None
    if emit_planes:
        if out_planes is not None:  # caller-provided (e.g. in place over the residual planes)
            oh, ol = out_planes
        else:
            oh = torch.empty(B, T, out_cols, dtype=torch.float16, device=xh.device)
            ol = torch.empty(B, T, out_cols, dtype=torch.float16, device=xh.device)
        d.out_hi = oh.data_ptr(); d.out_lo = ol.data_ptr(); d.out_plane_bs = oh.stride(0); d.out_plane_ld = oh.stride(1)
        d.out_plane_add = None if plane_add is None else plane_add.data_ptr()
        planes = (oh, ol)
    if _desc_only:
        return d, out, planes, (xh, xl, wh, wl)
    _abi.check(_abi.lib().pttspp_conv1d_cl(C.byref(d), _abi.stream_ptr(xh.device)))
    return out, planes


def aa_conv1d_cl(x, log_alpha, up_filter, down_filter, w_split, bias=None, K=1, dil=1, pad=0, res=None, beta=0.0,
                 out=None, out_div=0.0):
    """AA-Snake fused into the consuming conv (pttspp_aa_conv1d_cl): x [B, T, C] fp32 pre-activation, C in {32, 64}."""
    _abi.require_cuda(x, "aa_conv1d_cl")
    B, T, Cc = x.shape
    d, out, _, keep = conv1d_umma_cl((x, x), w_split, Cc, bias=bias, K=K, dil=dil, pad=pad, res=res, beta=beta, out=out,
                                     out_div=out_div, _desc_only=True)
    d.in_hi = None; d.in_lo = None
    d.in_ = x.data_ptr()
    d.in_bs = x.stride(0); d.in_ld = x.stride(1)
    _abi.check(_abi.lib().pttspp_aa_conv1d_cl(C.byref(d), _abi.ptr(log_alpha), _abi.ptr(up_filter), _abi.ptr(down_filter),
                                              _abi.stream_ptr(x.device)))
    return out


def conv1d_umma_dual_cl(x_planes, w_split, cout1, kw1, kw2):
    """One tcgen05 launch, two epilogues: columns [0, cout1) use kw1, the rest kw2 (kwargs of conv1d_umma_cl).
    w_split packs all Cout rows; returns ((out1, planes1), (out2, planes2))."""
    wh, wl, sc = w_split
    Cin = wh.shape[-1]
    cout2 = wh.shape[1] - cout1
    d1, o1, p1, keep1 = conv1d_umma_cl(x_planes, (wh, wl, sc), cout1, _desc_only=True, **kw1)
    wh2, wl2 = wh.view(-1, Cin)[cout1:], wl.view(-1, Cin)[cout1:]
    d2, o2, p2, keep2 = conv1d_umma_cl(x_planes, (wh2, wl2, sc), cout2, _desc_only=True, **kw2)
    _abi.check(_abi.lib().pttspp_conv1d_dual_cl(C.byref(d1), C.byref(d2), _abi.stream_ptr(x_planes[0].device)))
    return (o1, p1), (o2, p2)


class DiffNetStack:
    """Fused DiffNet residual-layer stack (pttspp_diffnet_*; denoiser.py:69-83 for 256 channels, kernel 3, dilation <= 8).

    layers: list of dicts with `dilated` [512, 256, 3], `dilated_bias` [512], `outp` [512, 256(, 1)], `outp_bias` [512]
    (torch layouts, reference channel order) and `dil`.  Packing (gate/filter interleave, split-fp16 planes) happens here.
    """

    def __init__(self, layers, device):
        self.device = torch.device(device)
        self._keep = []
        arr = (_abi.DiffNetLayer * len(layers))()
        for i, l in enumerate(layers):
            wdh, wdl, sd = pack_conv_weight_split(l["dilated"], interleave_halves=True, device=self.device)
            woh, wol, so = pack_conv_weight_split(l["outp"], device=self.device)
            bd = l["dilated_bias"].detach().float().cpu()
            half = bd.numel() // 2
            bdi = torch.stack([bd[:half], bd[half:]], dim=1).reshape(-1).contiguous().to(self.device)
            bo = l["outp_bias"].detach().float().contiguous().to(self.device)
            self._keep += [wdh, wdl, woh, wol, bdi, bo]
            arr[i].wd_hi, arr[i].wd_lo, arr[i].wo_hi, arr[i].wo_lo = (t.data_ptr() for t in (wdh, wdl, woh, wol))
            arr[i].bias_d, arr[i].bias_o = bdi.data_ptr(), bo.data_ptr()
            arr[i].scale_d, arr[i].scale_o, arr[i].dil = sd, so, int(l["dil"])
        self.n_layers = len(layers)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _abi.check(_abi.lib().pttspp_diffnet_create(arr, len(layers), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            _abi.lib().pttspp_diffnet_destroy(self.h)
            self.h = None

    def new_flags(self, B, T):
        n = _abi.lib().pttspp_diffnet_flags_bytes(self.h, B, T)
        return torch.zeros(n // 4, dtype=torch.int32, device=self.device)

    @staticmethod
    def interleave_cond(cond):
        """[..., 512] reference order (gate half | filter half) -> the kernel's interleaved column order."""
        half = cond.shape[-1] // 2
        return torch.stack([cond[..., :half], cond[..., half:]], dim=-1).reshape(cond.shape).contiguous()

    def run(self, cond, step_emb, y, skip, done, epoch, layer_begin=0, layer_end=None, skip_planes=None, dbg_z=None,
            dbg_prof=None):
        """cond [L, B, T, 512] (interleaved), step_emb [L + 1, 256], y = ((hi0, lo0), (hi1, lo1)) planes [B, T, 256]."""
        (h0, l0), (h1, l1) = y
        B, T, _ = h0.shape
        r = _abi.DiffNetRunDesc()
        r.B, r.T = B, T
        r.layer_begin, r.layer_end = layer_begin, self.n_layers if layer_end is None else layer_end
        r.cond, r.step_emb = cond.data_ptr(), step_emb.data_ptr()
        r.y_hi[0], r.y_lo[0], r.y_hi[1], r.y_lo[1] = h0.data_ptr(), l0.data_ptr(), h1.data_ptr(), l1.data_ptr()
        r.skip = skip.data_ptr()
        if skip_planes is not None:
            r.skip_hi, r.skip_lo = skip_planes[0].data_ptr(), skip_planes[1].data_ptr()
        r.done, r.epoch = done.data_ptr(), int(epoch)
        r.dbg_z = None if dbg_z is None else dbg_z.data_ptr()
        r.dbg_prof = None if dbg_prof is None else dbg_prof.data_ptr()
        with torch.cuda.device(self.device):
            _abi.check(_abi.lib().pttspp_diffnet_run(self.h, C.byref(r), _abi.stream_ptr(self.device)))


def philox_normal(numel, seed, offset, device="cuda"):
    """What `torch.empty(numel, device=device).normal_()` writes when the CUDA generator holds (seed, offset); returns
    (tensor, offset advance).  csrc/philox.cu."""
    out = torch.empty(int(numel), dtype=torch.float32, device=device)
    adv = C.c_uint64(0)
    with torch.cuda.device(out.device):
        _abi.check(_abi.lib().pttspp_philox_normal(_abi.ptr(out), int(numel), C.c_uint64(int(seed)), C.c_uint64(int(offset)),
                                                   C.byref(adv), _abi.stream_ptr(out.device)))
    return out, int(adv.value)
