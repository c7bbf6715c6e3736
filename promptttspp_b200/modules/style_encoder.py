"""Reference-mel style encoder.

Reference: promptttspp/modules/style_encoder.py:23-171 and promptttspp/modules/reference_encoder.py:21-124.
The sub-modules hold the parameters under the reference's state_dict keys; `StyleEncoder.forward` runs the
arithmetic through three C-ABI ops (csrc/style_encoder.cu): pttspp_conv2d_bn_relu x 6, pttspp_gru_last_state,
pttspp_style_token_attention.  CUDA tensors only.
"""
import math
from typing import Sequence

import torch
from torch import nn

from .. import _abi


class _TokenAttention(nn.Module):
    def __init__(self, q_dim, k_dim, v_dim, n_head, n_feat):
        super().__init__()
        self.linear_q = nn.Linear(q_dim, n_feat)
        self.linear_k = nn.Linear(k_dim, n_feat)
        self.linear_v = nn.Linear(v_dim, n_feat)
        self.linear_out = nn.Linear(n_feat, n_feat)


class StyleTokenLayer(nn.Module):
    def __init__(self, ref_embed_dim=128, gst_tokens=10, gst_token_dim=256, gst_heads=4):
        super().__init__()
        self.gst_embs = nn.Parameter(torch.randn(gst_tokens, gst_token_dim // gst_heads))
        d = gst_token_dim // gst_heads
        self.mha = _TokenAttention(ref_embed_dim, d, d, gst_heads, gst_token_dim)


class ReferenceEncoder(nn.Module):
    def __init__(self, idim=80, conv_layers=6, conv_chans_list: Sequence[int] = (32, 32, 64, 64, 128, 128),
                 conv_kernel_size=3, conv_stride=2, gru_layers=1, gru_units=128):
        super().__init__()
        assert conv_kernel_size % 2 == 1 and len(conv_chans_list) == conv_layers
        pad = (conv_kernel_size - 1) // 2
        mods, width = [], idim
        for i, out_ch in enumerate(conv_chans_list):
            in_ch = 1 if i == 0 else conv_chans_list[i - 1]
            mods += [nn.Conv2d(in_ch, out_ch, conv_kernel_size, stride=conv_stride, padding=pad, bias=False),
                     nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True)]
            width = (width - conv_kernel_size + 2 * pad) // conv_stride + 1
        self.convs = nn.Sequential(*mods)
        self.gru = nn.GRU(width * conv_chans_list[-1], gru_units, gru_layers, batch_first=True)


class StyleEncoder(nn.Module):
    def __init__(self, idim=80, gst_tokens=10, gst_token_dim=256, gst_heads=4, conv_layers=6,
                 conv_chans_list: Sequence[int] = (32, 32, 64, 64, 128, 128), conv_kernel_size=3,
                 conv_stride=2, gru_layers=1, gru_units=128):
        super().__init__()
        self.ref_enc = ReferenceEncoder(idim, conv_layers, conv_chans_list, conv_kernel_size,
                                        conv_stride, gru_layers, gru_units)
        self.stl = StyleTokenLayer(gru_units, gst_tokens, gst_token_dim, gst_heads)

    @torch.no_grad()
    def forward(self, speech, in_lens=None):
        """speech: [B, idim, Lmax] normalised mel (CUDA), in_lens: [B] frame counts or None -> [B, token_dim, 1]."""
        _abi.require_cuda(speech, "StyleEncoder.forward")
        lib = _abi.lib()
        dev = speech.device
        enc, stl = self.ref_enc, self.stl
        x = speech.float().transpose(1, 2).unsqueeze(1).contiguous()  # (B, 1, Lmax, idim), reference_encoder.py:103
        B = x.shape[0]
        with torch.cuda.device(dev):
            stream = _abi.stream_ptr(dev)
            mods = list(enc.convs)
            for i in range(0, len(mods), 3):
                conv, bn = mods[i], mods[i + 1]
                k, st, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
                _, cin, H, W = x.shape
                Ho, Wo = (H + 2 * pad - k) // st + 1, (W + 2 * pad - k) // st + 1
                # eval-mode BatchNorm2d folded to y * scale + shift
                scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
                shift = (bn.bias - bn.running_mean * scale).float().contiguous()
                w = conv.weight.detach().float().contiguous()
                out = torch.empty(B, conv.out_channels, Ho, Wo, device=dev)
                _abi.check(lib.pttspp_conv2d_bn_relu(_abi.ptr(x), _abi.ptr(w), _abi.ptr(scale), _abi.ptr(shift), _abi.ptr(out),
                                                     B, cin, H, W, conv.out_channels, k, st, pad, stream))
                x = out
            hs = x.transpose(1, 2).contiguous().view(B, x.shape[2], -1)  # (B, Lmax', C * idim'), :105-108
            T, I = hs.shape[1], hs.shape[2]
            lens = None
            if in_lens is not None:  # :113-118, without the host round trip of hs_lens.to("cpu")
                n_sub = len(mods) // 3
                lens = torch.ceil(in_lens.to(dev).float() / (mods[0].stride[0] ** n_sub)).long().clamp(min=1).contiguous()
            gru = enc.gru
            Hn = gru.hidden_size
            ref = torch.empty(B, Hn, device=dev)
            _abi.check(lib.pttspp_gru_last_state(
                _abi.ptr(hs), None if lens is None else _abi.ptr(lens), B, T, I, Hn,
                _abi.ptr(gru.weight_ih_l0.detach().float().contiguous()), _abi.ptr(gru.weight_hh_l0.detach().float().contiguous()),
                _abi.ptr(gru.bias_ih_l0.detach().float().contiguous()), _abi.ptr(gru.bias_hh_l0.detach().float().contiguous()),
                _abi.ptr(ref), stream))
            mha = stl.mha
            Fd = mha.linear_q.out_features
            Tk, Dk = stl.gst_embs.shape
            heads = Fd // Dk
            out = torch.empty(B, Fd, device=dev)
            f = lambda t: _abi.ptr(t.detach().float().contiguous())
            keep = [t.detach().float().contiguous() for t in (
                stl.gst_embs, mha.linear_q.weight, mha.linear_q.bias, mha.linear_k.weight, mha.linear_k.bias,
                mha.linear_v.weight, mha.linear_v.bias, mha.linear_out.weight, mha.linear_out.bias)]
            _abi.check(lib.pttspp_style_token_attention(_abi.ptr(ref), B, Hn, _abi.ptr(keep[0]), Tk, Dk, heads, Fd,
                                                        *[_abi.ptr(t) for t in keep[1:]], _abi.ptr(out), stream))
        return out.unsqueeze(-1)
