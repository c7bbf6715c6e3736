"""DiffNet (WaveNet-style denoiser) parameter holder.

Reference: promptttspp/modules/denoiser.py:50-143.  Forward = csrc/acoustic.cu
(diffnet_step); the step-embedding MLP is folded into a [K_step, layers, C]
table at weight-pack time because it depends only on the integer step.
"""
from torch import nn


def _kaiming_conv1d(*args, **kwargs):
    layer = nn.Conv1d(*args, **kwargs)
    nn.init.kaiming_normal_(layer.weight)
    return layer


class ResidualBlock(nn.Module):
    def __init__(self, encoder_hidden, residual_channels, kernel_size, dilation):
        super().__init__()
        self.dilation = dilation
        self.dilated_conv = _kaiming_conv1d(
            residual_channels, 2 * residual_channels, kernel_size,
            padding=(kernel_size * dilation - dilation) // 2, dilation=dilation,
        )
        self.diffusion_projection = nn.Linear(residual_channels, residual_channels)
        self.conditioner_projection = _kaiming_conv1d(encoder_hidden, 2 * residual_channels, 1)
        self.output_projection = _kaiming_conv1d(residual_channels, 2 * residual_channels, 1)


class _MishMarker(nn.Module):
    pass


class DiffNet(nn.Module):
    def __init__(self, in_dim=80, encoder_hidden_dim=256, residual_layers=20,
                 residual_channels=256, kernel_size=3, dilation_cycle_length=4, scale=1):
        super().__init__()
        self.in_dim = in_dim
        self.encoder_hidden_dim = encoder_hidden_dim
        self.residual_channels = residual_channels
        self.kernel_size = kernel_size
        self.dilation_cycle_length = dilation_cycle_length
        self.scale = scale
        self.input_projection = _kaiming_conv1d(in_dim, residual_channels, 1)
        dim = residual_channels
        # index 1 is the parameter-free Mish, so the keys are mlp.0.* / mlp.2.*
        self.mlp = nn.Sequential(nn.Linear(dim, dim * 4), _MishMarker(), nn.Linear(dim * 4, dim))
        self.residual_layers = nn.ModuleList(
            [ResidualBlock(encoder_hidden_dim, residual_channels, kernel_size,
                           2 ** (i % dilation_cycle_length)) for i in range(residual_layers)]
        )
        self.skip_projection = _kaiming_conv1d(residual_channels, residual_channels, 1)
        self.output_projection = _kaiming_conv1d(residual_channels, in_dim, 1)
        nn.init.zeros_(self.output_projection.weight)
