"""Conformer text encoder -- parameter holders with the reference's key names.

Reference: promptttspp/modules/esp/__init__.py:11-65 (wrapper),
esp/conformer/encoder.py:60-282, esp/conformer/encoder_layer.py:15-162,
esp/conformer/convolution.py:13-85, esp/transformer/attention.py:15-305,
esp/transformer/multi_layer_conv.py:12-67, esp/transformer/layer_norm.py:21.

Only the configuration family that the shipped yaml files use is supported
(macaron conv1d feed-forward, rel-pos self-attention in its ``legacy`` or
``new`` flavour, Swish conv module); anything else raises at construction.
The forward pass is csrc/acoustic.cu::encoder_forward.
"""
import torch
from torch import nn


class _RelPosSelfAttention(nn.Module):
    """linear_{q,k,v,out,pos} + pos_bias_{u,v} (attention.py:114-141 / 209-236)."""

    def __init__(self, n_head, n_feat):
        super().__init__()
        assert n_feat % n_head == 0
        self.h, self.d_k = n_head, n_feat // n_head
        self.linear_q = nn.Linear(n_feat, n_feat)
        self.linear_k = nn.Linear(n_feat, n_feat)
        self.linear_v = nn.Linear(n_feat, n_feat)
        self.linear_out = nn.Linear(n_feat, n_feat)
        self.linear_pos = nn.Linear(n_feat, n_feat, bias=False)
        self.pos_bias_u = nn.Parameter(torch.empty(self.h, self.d_k))
        self.pos_bias_v = nn.Parameter(torch.empty(self.h, self.d_k))
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)


class _ConvFeedForward(nn.Module):
    """w_1 / w_2 of MultiLayeredConv1d (multi_layer_conv.py:26-50)."""

    def __init__(self, chans, hidden, kernel_size):
        super().__init__()
        pad = (kernel_size - 1) // 2
        self.w_1 = nn.Conv1d(chans, hidden, kernel_size, padding=pad)
        self.w_2 = nn.Conv1d(hidden, chans, kernel_size, padding=pad)


class _ConvModule(nn.Module):
    """pointwise/GLU/depthwise/BatchNorm/pointwise (convolution.py:22-56)."""

    def __init__(self, chans, kernel_size):
        super().__init__()
        assert (kernel_size - 1) % 2 == 0
        self.pointwise_conv1 = nn.Conv1d(chans, 2 * chans, 1)
        self.depthwise_conv = nn.Conv1d(
            chans, chans, kernel_size, padding=(kernel_size - 1) // 2, groups=chans
        )
        self.norm = nn.BatchNorm1d(chans)
        self.pointwise_conv2 = nn.Conv1d(chans, chans, 1)


class _Block(nn.Module):
    def __init__(self, size, heads, hidden, ff_kernel, cnn_kernel):
        super().__init__()
        self.self_attn = _RelPosSelfAttention(heads, size)
        self.feed_forward = _ConvFeedForward(size, hidden, ff_kernel)
        self.feed_forward_macaron = _ConvFeedForward(size, hidden, ff_kernel)
        self.conv_module = _ConvModule(size, cnn_kernel)
        for name in ("norm_ff", "norm_mha", "norm_ff_macaron", "norm_conv", "norm_final"):
            setattr(self, name, nn.LayerNorm(size, eps=1e-12))


class _Encoder(nn.Module):
    def __init__(self, size, heads, hidden, blocks, ff_kernel, cnn_kernel):
        super().__init__()
        self.encoders = nn.Sequential(
            *[_Block(size, heads, hidden, ff_kernel, cnn_kernel) for _ in range(blocks)]
        )
        self.after_norm = nn.LayerNorm(size, eps=1e-12)


class ConformerEncoder(nn.Module):
    def __init__(
        self,
        idim=8,
        attention_dim=8,
        return_mask=False,
        rel_pos_type=None,
        attention_heads=4,
        linear_units=2048,
        num_blocks=6,
        dropout_rate=0.1,
        positionwise_layer_type="linear",
        positionwise_conv_kernel_size=1,
        pos_enc_layer_type="abs_pos",
        selfattention_layer_type="selfattn",
        activation_type="swish",
        macaron_style=False,
        use_cnn_module=False,
        cnn_module_kernel=31,
        **unused,
    ):
        super().__init__()
        if rel_pos_type is None or rel_pos_type == "legacy":
            self.rel_pos_type = "legacy"
        elif rel_pos_type == "new":
            self.rel_pos_type = "new"
        else:
            raise ValueError(f"Unknown relative positional encoding type: {rel_pos_type}")
        unsupported = []
        if idim != attention_dim:
            unsupported.append("idim != attention_dim")
        if positionwise_layer_type != "conv1d":
            unsupported.append(f"positionwise_layer_type={positionwise_layer_type}")
        if pos_enc_layer_type != "rel_pos" or selfattention_layer_type != "rel_selfattn":
            unsupported.append("non rel-pos attention")
        if activation_type != "swish" or not macaron_style or not use_cnn_module:
            unsupported.append("non-macaron / non-swish / no cnn module")
        if return_mask:
            unsupported.append("return_mask=True")
        if unsupported:
            raise NotImplementedError(
                "promptttspp_b200 ConformerEncoder supports the shipped "
                "prompttts_mdn_v2_wo_erg_final[_demo].yaml family only: " + ", ".join(unsupported)
            )
        self._out_dim = attention_dim
        self.heads = attention_heads
        self.linear_units = linear_units
        self.ff_kernel = positionwise_conv_kernel_size
        self.cnn_kernel = cnn_module_kernel
        self.num_blocks = num_blocks
        self.encoder = _Encoder(
            attention_dim, attention_heads, linear_units, num_blocks,
            positionwise_conv_kernel_size, cnn_module_kernel,
        )

    @property
    def out_dim(self):
        return self._out_dim
