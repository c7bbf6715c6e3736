"""Reference-mel style encoder -- parameter holder only ("next").

Reference: promptttspp/modules/style_encoder.py:23-171 and
promptttspp/modules/reference_encoder.py:21-124.  The keys are declared so a
reference checkpoint loads with ``strict=True``; the ``reference_mel=`` style
path itself is not on the accelerated hot path yet and raises when used.
"""
from typing import Sequence

import torch
from torch import nn


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

    def forward(self, speech, in_lens=None):
        raise NotImplementedError(
            "the reference_mel style path (SURVEY.md 8f3) is not accelerated yet; "
            "use style_prompt=..."
        )
