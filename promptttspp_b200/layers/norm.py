"""Channel LayerNorm parameter holder (reference: promptttspp/layers/norm.py:19-32).

Only the parameters live here (``gamma``/``beta`` of shape [1, C, 1]); the
arithmetic runs in csrc/norm.cu (pttspp_layernorm_cl).
"""
import torch
from torch import nn


class LayerNorm(nn.Module):
    def __init__(self, channels, eps=1e-5):
        super().__init__()
        self.channels = channels
        self.eps = eps
        self.gamma = nn.Parameter(torch.ones(1, channels, 1))
        self.beta = nn.Parameter(torch.zeros(1, channels, 1))
