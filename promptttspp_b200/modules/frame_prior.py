"""Frame-prior network parameter holder (reference: promptttspp/modules/frame_prior.py:23-92)."""
import torch
from torch import nn


class LayerNorm(nn.Module):
    """gamma/beta of shape [C] (frame_prior.py:23-36)."""

    def __init__(self, channels, eps=1e-5):
        super().__init__()
        self.channels, self.eps = channels, eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))


class FramePriorNetwork(nn.Module):
    def __init__(self, out_channels, hidden_channels, n_layers, kernel_size, p_dropout,
                 pos_enc_p_dropout=0.1, use_pos_enc=True, use_rel=False):
        super().__init__()
        if not use_pos_enc or use_rel:
            raise NotImplementedError("only use_pos_enc=True, use_rel=False is supported")
        self.out_channels, self.hidden_channels = out_channels, hidden_channels
        self.n_layers, self.kernel_size = n_layers, kernel_size
        self.norm_emb = LayerNorm(hidden_channels)
        self.convs = nn.ModuleList(
            [nn.Conv1d(hidden_channels, hidden_channels, kernel_size, padding=kernel_size // 2)
             for _ in range(n_layers)]
        )
        self.norms = nn.ModuleList([LayerNorm(hidden_channels) for _ in range(n_layers)])
