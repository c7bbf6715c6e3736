"""Duration / pitch predictors and the variance adaptor -- parameter holders.

Reference: promptttspp/modules/variance_adaptor.py:23-206.  The inference
arithmetic (conv -> ReLU -> channel-LN -> mask stacks, MDN head, duration
quantisation, length regulator, frame prior, pitch embedding) is
csrc/acoustic.cu.
"""
from torch import nn

from ..layers.norm import LayerNorm
from .mdn import MDNLayer


class PredictorLayer(nn.Module):
    def __init__(self, channels, kernel_size, dropout):
        super().__init__()
        self.conv = nn.Conv1d(channels, channels, kernel_size, padding=kernel_size // 2)
        self.norm = LayerNorm(channels)


class Predictor(nn.Module):
    def __init__(self, channels, out_channels, kernel_size, dropout, num_layers, detach=False):
        super().__init__()
        self.kernel_size = kernel_size
        self.layers = nn.ModuleList(
            [PredictorLayer(channels, kernel_size, dropout) for _ in range(num_layers)]
        )
        self.out_layer = nn.Conv1d(channels, out_channels, 1)
        self.detach = detach


class MDNPredictor(nn.Module):
    def __init__(self, channels, out_channels, kernel_size, dropout, num_layers,
                 num_gaussians=4, dim_wise=True, detach=False, disable_amp=False):
        super().__init__()
        self.kernel_size = kernel_size
        self.layers = nn.ModuleList(
            [PredictorLayer(channels, kernel_size, dropout) for _ in range(num_layers)]
        )
        self.out_layer = MDNLayer(channels, out_channels, num_gaussians, dim_wise)
        self.detach = detach
        self.disable_amp = disable_amp


class VarianceAdaptor(nn.Module):
    def __init__(self, duration_predictor, pitch_predictor, pitch_emb,
                 energy_predictor=None, energy_emb=None, frame_prior_network=None):
        super().__init__()
        if energy_predictor is not None or energy_emb is not None:
            raise NotImplementedError("energy predictor is not part of the shipped configs")
        self.duration_predictor = duration_predictor
        self.pitch_predictor = pitch_predictor
        self.pitch_emb = pitch_emb
        self.energy_predictor = None
        self.energy_emb = None
        self.frame_prior_network = frame_prior_network
