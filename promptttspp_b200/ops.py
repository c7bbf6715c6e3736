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


def aa_snake_cl(x, log_alpha, up_filter, down_filter):
    _abi.require_cuda(x, "aa_snake_cl")
    B, L, Cc = x.shape
    y = torch.empty_like(x)
    _abi.check(_abi.lib().pttspp_aa_snake_cl(_abi.ptr(x), _abi.ptr(y), B, L, Cc, _abi.ptr(log_alpha),
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
