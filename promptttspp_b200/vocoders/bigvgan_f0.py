"""F0-aware BigVGAN -- drop-in for promptttspp.vocoders.F0AwareBigVGAN (the vocoder app.py and
egs/proposed/bin/synthesize.py instantiate by default, conf/vocoder/bigvgan_f0.yaml).

Same constructor kwargs, same state_dict keys as the reference (promptttspp/vocoders/bigvgan_f0.py:25-96: the BigVGAN
keys plus `m_source.l_linear.{weight,bias}` and `noise_convs.{i}.{weight,bias}`), same
`forward(mel[B, 80, T], f0[B, 1, T]) -> wav[B, 1, 240*T]` (:98-115).  The arithmetic is pttspp_nsf_source +
pttspp_bigvgan_forward_f0 (csrc/bigvgan.cu).
"""
import ctypes as C

import torch
from torch import nn

from .. import _abi
from .bigvgan import BigVGAN
from .nsf import SourceModuleHnNSF


class F0AwareBigVGAN(BigVGAN):
    def __init__(self, sampling_rate, harmonic_num, in_channel, upsample_initial_channel, upsample_rates,
                 upsample_kernel_sizes, resblock_kernel_sizes, resblock_dilations):
        super().__init__(in_channel, upsample_initial_channel, upsample_rates, upsample_kernel_sizes,
                         resblock_kernel_sizes, resblock_dilations)
        self.m_source = SourceModuleHnNSF(sampling_rate=sampling_rate, harmonic_num=harmonic_num)
        self.noise_convs = nn.ModuleList()
        rates = self.upsample_rates
        for i in range(len(rates)):
            ch = upsample_initial_channel // (2 ** (i + 1))
            if i + 1 < len(rates):
                s = 1
                for r in rates[i + 1:]:
                    s *= r
                self.noise_convs.append(nn.Conv1d(1, ch, kernel_size=s * 2, stride=s, padding=s // 2))
            else:
                self.noise_convs.append(nn.Conv1d(1, ch, 1))

    @torch.no_grad()
    def forward(self, x, f0, source_noise=None):
        """x: mel [B, in_channel, T]; f0: [B, 1, T] in Hz (0 on unvoiced frames) -> waveform [B, 1, T * hop].
        `source_noise` (vocoders.nsf.SourceNoise) injects the reference's random draws (parity tests); by default they
        are drawn from torch's CUDA generator in the reference's order."""
        _abi.require_cuda(x, "F0AwareBigVGAN.forward")
        if x.dim() != 3 or x.shape[1] != self.in_channel:
            raise ValueError(f"expected mel of shape [B, {self.in_channel}, T], got {tuple(x.shape)}")
        if f0.dim() != 3 or f0.shape[1] != 1 or f0.shape[0] != x.shape[0] or f0.shape[2] != x.shape[2]:
            raise ValueError(f"expected f0 of shape [B, 1, T] = [{x.shape[0]}, 1, {x.shape[2]}], got {tuple(f0.shape)}")
        x = x.contiguous().float()
        B, _, T = x.shape
        out = torch.empty(B, 1, T * self.hop, dtype=torch.float32, device=x.device)
        if B == 0 or T == 0:
            return out
        har = self.m_source(f0.to(x.device)[:, 0, :], self.hop, source_noise)
        with torch.cuda.device(x.device):
            nat = self._handle(x.device)
            lib = _abi.lib()
            nbytes = lib.pttspp_bigvgan_workspace_bytes(nat.h, B, T)
            ws = nat.workspace(nbytes, x.device)
            _abi.check(lib.pttspp_bigvgan_forward_f0(nat.h, _abi.ptr(x), _abi.ptr(har), B, T, _abi.ptr(out), _abi.ptr(ws),
                                                     C.c_size_t(ws.numel()), _abi.stream_ptr(x.device)))
        return out
