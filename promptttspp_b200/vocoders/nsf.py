"""Harmonic-plus-noise source of the F0-aware vocoder -- parameter holder + native call.

Mirrors promptttspp/vocoders/nsf.py: `SourceModuleHnNSF(sampling_rate, harmonic_num, sine_amp=0.1,
add_noise_std=0.003, voiced_threshod=0)` with the state_dict key `l_linear.{weight,bias}`; the arithmetic
(SineGen's phase accumulation, the voiced/unvoiced mix, Linear + tanh) is pttspp_nsf_source (csrc/bigvgan.cu).
"""
import ctypes as C
from typing import NamedTuple, Optional

import torch
from torch import nn

from .. import _abi


class SourceNoise(NamedTuple):
    """The reference's random draws, in its call order (nsf.py:64 torch.rand, :143 randn_like, :205 randn_like)."""
    rand_ini: torch.Tensor   # [B, H], column 0 is forced to 0 like nsf.py:67
    noise: torch.Tensor      # [B, L, H]
    noise_uv: Optional[torch.Tensor] = None  # [B, L, 1]; drawn only to keep the generator in step, unused by the vocoder


class SourceModuleHnNSF(nn.Module):
    def __init__(self, sampling_rate, harmonic_num=0, sine_amp=0.1, add_noise_std=0.003, voiced_threshod=0):
        super().__init__()
        self.sampling_rate = float(sampling_rate)
        self.harmonic_num = int(harmonic_num)
        self.sine_amp = float(sine_amp)
        self.noise_std = float(add_noise_std)
        self.voiced_threshold = float(voiced_threshod)
        self.l_linear = nn.Linear(self.harmonic_num + 1, 1)
        self._ws = None

    def draw_noise(self, B, L, device):
        """Same shapes, order and generator (torch's current CUDA generator) as the reference's forward."""
        H = self.harmonic_num + 1
        rand_ini = torch.rand(B, H, device=device)
        rand_ini[:, 0] = 0
        noise = torch.randn(B, L, H, device=device)
        noise_uv = torch.randn(B, L, 1, device=device)
        return SourceNoise(rand_ini, noise, noise_uv)

    @torch.no_grad()
    def forward(self, f0, hop, noise: Optional[SourceNoise] = None):
        """f0: [B, T] (Hz, 0 = unvoiced, frame rate) -> har_source [B, T*hop] (the reference returns it as [B, L, 1])."""
        _abi.require_cuda(f0, "SourceModuleHnNSF.forward")
        f0 = f0.contiguous().float()
        B, T = f0.shape
        L = T * int(hop)
        H = self.harmonic_num + 1
        if noise is None:
            noise = self.draw_noise(B, L, f0.device)
        rand_ini = noise.rand_ini.to(f0.device).float().contiguous().clone()
        rand_ini[:, 0] = 0
        nz = noise.noise.to(f0.device).float().contiguous()
        if tuple(rand_ini.shape) != (B, H) or tuple(nz.shape) != (B, L, H):
            raise ValueError(f"source noise shapes {tuple(rand_ini.shape)}, {tuple(nz.shape)} != ({B},{H}), ({B},{L},{H})")
        out = torch.empty(B, L, dtype=torch.float32, device=f0.device)
        if B == 0 or L == 0:
            return out
        lib = _abi.lib()
        w = self.l_linear.weight.detach().float().contiguous().view(-1)
        b = self.l_linear.bias.detach().float().contiguous()
        with torch.cuda.device(f0.device):
            nbytes = lib.pttspp_nsf_source_workspace_bytes(B, T, int(hop), self.harmonic_num)
            if self._ws is None or self._ws.numel() < nbytes or self._ws.device != f0.device:
                self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=f0.device)
            _abi.check(lib.pttspp_nsf_source(_abi.ptr(f0), B, T, int(hop), self.sampling_rate, self.harmonic_num,
                                             self.sine_amp, self.noise_std, self.voiced_threshold, _abi.ptr(rand_ini),
                                             _abi.ptr(nz), _abi.ptr(w), _abi.ptr(b), _abi.ptr(out), _abi.ptr(self._ws),
                                             C.c_size_t(self._ws.numel()), _abi.stream_ptr(f0.device)))
        return out
