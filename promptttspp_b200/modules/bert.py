"""BERT encoder for the prompt encoder's sentence embedding.

The reference calls HF transformers' `BertModel` (promptttspp/modules/prompt_encoder.py:22-38: tokenizer ->
`self.model(**inputs).last_hidden_state[:, 0, :]`).  `NativeBert` holds the parameters under the HF state_dict names
(so `prompt_encoder.bert.model.*` checkpoint keys load) and runs the forward through the C ABI: pttspp_bert_embed,
pttspp_layernorm_cl, pttspp_conv1d_cl (Linear / GELU, fp32 CUDA cores), pttspp_mha_masked.  Third-party arithmetic: the
algorithm restated is `modeling_bert.py` of transformers 5.5.0 (BertEmbeddings, BertSelfAttention eager path, BertSelfOutput,
BertIntermediate with exact-erf GELU, BertOutput); it is pinned against HF's own module in tests/golden/bert_small.npz.
"""
import ctypes as C

import torch
from torch import nn

from .. import _abi, ops


class _SelfAttention(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)


class _SelfOutput(nn.Module):
    def __init__(self, i, h, eps):
        super().__init__()
        self.dense = nn.Linear(i, h)
        self.LayerNorm = nn.LayerNorm(h, eps=eps)


class _Attention(nn.Module):
    def __init__(self, h, eps):
        super().__init__()
        self.self = _SelfAttention(h)
        self.output = _SelfOutput(h, h, eps)


class _Intermediate(nn.Module):
    def __init__(self, h, i):
        super().__init__()
        self.dense = nn.Linear(h, i)


class _Layer(nn.Module):
    def __init__(self, h, i, eps):
        super().__init__()
        self.attention = _Attention(h, eps)
        self.intermediate = _Intermediate(h, i)
        self.output = _SelfOutput(i, h, eps)


class _Encoder(nn.Module):
    def __init__(self, n, h, i, eps):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(h, i, eps) for _ in range(n)])


class _Embeddings(nn.Module):
    def __init__(self, vocab, h, max_pos, types, eps):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, h, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, h)
        self.token_type_embeddings = nn.Embedding(types, h)
        self.LayerNorm = nn.LayerNorm(h, eps=eps)


class _Pooler(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.dense = nn.Linear(h, h)  # kept for the checkpoint keys; the reference reads last_hidden_state[:, 0]


class NativeBert(nn.Module):
    """HF `BertModel` parameter tree + native forward.  Defaults = bert-base-uncased's config."""

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12):
        super().__init__()
        assert hidden_size % num_attention_heads == 0 and hidden_size % 16 == 0 and intermediate_size % 16 == 0
        self.heads, self.eps = num_attention_heads, layer_norm_eps
        self.embeddings = _Embeddings(vocab_size, hidden_size, max_position_embeddings, type_vocab_size, layer_norm_eps)
        self.encoder = _Encoder(num_hidden_layers, hidden_size, intermediate_size, layer_norm_eps)
        self.pooler = _Pooler(hidden_size)
        self._packed, self._sig = None, None

    def weights_signature(self):
        return tuple((k, v.data_ptr(), v._version, str(v.device)) for k, v in self.state_dict(keep_vars=True).items())

    def _pack(self, device):
        sig = self.weights_signature()
        if self._packed is not None and sig == self._sig:
            return self._packed
        f = lambda t: t.detach().float().to(device).contiguous()
        layers = []
        for l in self.encoder.layer:
            a = l.attention
            wqkv = torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0)
            bqkv = torch.cat([a.self.query.bias, a.self.key.bias, a.self.value.bias], 0)
            layers.append(dict(
                wqkv=ops.pack_conv_weight(wqkv, device=device), bqkv=f(bqkv),
                wo=ops.pack_conv_weight(a.output.dense.weight, device=device), bo=f(a.output.dense.bias),
                g1=f(a.output.LayerNorm.weight), b1=f(a.output.LayerNorm.bias),
                wi=ops.pack_conv_weight(l.intermediate.dense.weight, device=device), bi=f(l.intermediate.dense.bias),
                wf=ops.pack_conv_weight(l.output.dense.weight, device=device), bf=f(l.output.dense.bias),
                g2=f(l.output.LayerNorm.weight), b2=f(l.output.LayerNorm.bias)))
        e = self.embeddings
        self._packed = dict(word=f(e.word_embeddings.weight), pos=f(e.position_embeddings.weight),
                            type0=f(e.token_type_embeddings.weight[0]), g=f(e.LayerNorm.weight), b=f(e.LayerNorm.bias),
                            layers=layers)
        self._sig = sig
        return self._packed

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None):
        """input_ids [B, T] int64 (CUDA), attention_mask [B, T] (1 = token) -> last_hidden_state [B, T, hidden]."""
        _abi.require_cuda(input_ids, "NativeBert.forward")
        dev = input_ids.device
        ids = input_ids.to(torch.int64).contiguous()
        B, T = ids.shape
        if T > self.embeddings.position_embeddings.num_embeddings:
            raise ValueError("sequence longer than max_position_embeddings")
        mask = None if attention_mask is None else attention_mask.to(device=dev, dtype=torch.int64).contiguous()
        p = self._pack(dev)
        lib = _abi.lib()
        Hd = p["word"].shape[1]
        dk = Hd // self.heads
        with torch.cuda.device(dev):
            st = _abi.stream_ptr(dev)
            x = torch.empty(B, T, Hd, device=dev)
            _abi.check(lib.pttspp_bert_embed(_abi.ptr(ids), B, T, _abi.ptr(p["word"]), p["word"].shape[0], _abi.ptr(p["pos"]),
                                             _abi.ptr(p["type0"]), Hd, _abi.ptr(x), st))
            x = ops.layernorm_cl(x, p["g"], p["b"], self.eps)
            for L in p["layers"]:
                qkv = ops.conv1d_cl(x, L["wqkv"], 3 * Hd, bias=L["bqkv"], impl=1)
                ctx = torch.empty(B, T, Hd, device=dev)
                _abi.check(lib.pttspp_mha_masked(_abi.ptr(qkv), None if mask is None else _abi.ptr(mask), B, T, self.heads, dk,
                                                 _abi.ptr(ctx), st))
                y = ops.conv1d_cl(ctx, L["wo"], Hd, bias=L["bo"], res=x, impl=1)          # dense + residual
                x = ops.layernorm_cl(y, L["g1"], L["b1"], self.eps)
                h = ops.conv1d_cl(x, L["wi"], L["bi"].numel(), bias=L["bi"], act=ops.ACT_GELU, impl=1)
                y = ops.conv1d_cl(h, L["wf"], Hd, bias=L["bf"], res=x, impl=1)
                x = ops.layernorm_cl(y, L["g2"], L["b2"], self.eps)
        return x
