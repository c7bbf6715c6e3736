"""Anti-aliased Snake activation parameter holders.

Reference: promptttspp/layers/activations.py:22-138.  The 2x up-sample ->
Snake -> 2x down-sample chain is one fused kernel (csrc/aa_snake.cu); these
classes only own ``alpha`` and the two 12-tap Kaiser-sinc filter buffers so the
``state_dict`` keys (``act.alpha``, ``up.filter``, ``down.lowpass.filter``)
match a reference checkpoint.
"""
import math

import torch
from torch import nn


def kaiser_sinc_filter1d(cutoff, half_width, kernel_size):
    """Kaiser-windowed sinc low-pass, normalised to unit DC gain -> [1, 1, K].

    Same design rule as the reference (activations.py:47-71): the Kaiser beta
    follows from the stop-band attenuation A = 2.285*(K/2-1)*pi*4*half_width+7.95.
    """
    half = kernel_size // 2
    atten = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    if atten > 50.0:
        beta = 0.1102 * (atten - 8.7)
    elif atten >= 21.0:
        beta = 0.5842 * (atten - 21) ** 0.4 + 0.07886 * (atten - 21.0)
    else:
        beta = 0.0
    win = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        n = torch.arange(-half, half) + 0.5
    else:
        n = torch.arange(kernel_size) - half
    if cutoff == 0:
        return torch.zeros(1, 1, kernel_size)
    taps = 2 * cutoff * win * torch.sinc(2 * cutoff * n)
    taps = taps / taps.sum()
    return taps.view(1, 1, kernel_size)


class UpSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.register_buffer(
            "filter", kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, self.kernel_size)
        )


class LowPassFilter1d(nn.Module):
    def __init__(self, cutoff=0.5, half_width=0.6, stride=1, kernel_size=12):
        super().__init__()
        if cutoff < -0.0:
            raise ValueError("Minimum cutoff must be larger than zero.")
        if cutoff > 0.5:
            raise ValueError("A cutoff above 0.5 does not make sense.")
        self.stride = stride
        self.kernel_size = kernel_size
        self.register_buffer("filter", kaiser_sinc_filter1d(cutoff, half_width, kernel_size))


class DownSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.lowpass = LowPassFilter1d(0.5 / ratio, 0.6 / ratio, ratio, self.kernel_size)


class Snake(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.alpha = nn.Parameter(torch.zeros(1, channels, 1))


class AntiAliasActivation(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.up = UpSample1d(2, 12)
        self.act = Snake(channels)
        self.down = DownSample1d(2, 12)
