"""MelSpectrogramTransform -- drop-in for promptttspp.transforms.MelSpectrogramTransform
(reference: promptttspp/transforms/mel.py:18-34, a torchaudio.transforms.MelSpectrogram subclass configured by
egs/proposed/bin/conf/transforms/mel.yaml; callers app.py:93-100, egs/proposed/bin/synthesize.py:109,172-175).

Same constructor keywords as torchaudio's MelSpectrogram, same `to_spec / spec_to_mel / to_mel / forward` methods and
`sample_rate` attribute.  The Hann window and the mel filterbank are built once on the host with torchaudio's fp32
recipes (`torch.hann_window`, `functional.melscale_fbanks`: restated below, torchaudio itself is not imported); the
transform itself is pttspp_mel_spectrogram (csrc/mel.cu).  CUDA tensors only -- no host fallback.
"""
import math

import torch
from torch import nn

from .. import _abi


def _hz_to_mel(freq: float, mel_scale: str) -> float:
    if mel_scale == "htk":
        return 2595.0 * math.log10(1.0 + (freq / 700.0))
    f_sp = 200.0 / 3
    mels = freq / f_sp
    min_log_hz = 1000.0
    if freq >= min_log_hz:
        mels = min_log_hz / f_sp + math.log(freq / min_log_hz) / (math.log(6.4) / 27.0)
    return mels


def _mel_to_hz(mels: torch.Tensor, mel_scale: str) -> torch.Tensor:
    if mel_scale == "htk":
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk"):
    """Triangular filterbank [n_freqs, n_mels], fp32 -- torchaudio.functional.melscale_fbanks (the MelScale buffer `fb`)."""
    if norm is not None and norm != "slaney":
        raise ValueError('norm must be one of None or "slaney"')
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale), n_mels + 2)
    f_pts = _mel_to_hz(m_pts, mel_scale)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2: n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)
    return fb


class MelSpectrogramTransform(nn.Module):
    def __init__(self, sample_rate=16000, n_fft=400, win_length=None, hop_length=None, f_min=0.0, f_max=None, pad=0,
                 n_mels=128, window_fn=torch.hann_window, power=2.0, normalized=False, wkwargs=None, center=True,
                 pad_mode="reflect", onesided=None, norm=None, mel_scale="htk"):
        super().__init__()
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.win_length = win_length if win_length is not None else n_fft
        self.hop_length = hop_length if hop_length is not None else self.win_length // 2
        self.n_mels = n_mels
        self.f_min = f_min
        self.f_max = f_max if f_max is not None else float(sample_rate // 2)
        self.power = power
        if pad != 0 or normalized or not center or pad_mode != "reflect" or onesided is False:
            raise NotImplementedError("MelSpectrogramTransform: only pad=0, normalized=False, center=True, "
                                      "pad_mode='reflect', one-sided spectra are supported (conf/transforms/mel.yaml)")
        if float(power) not in (1.0, 2.0):
            raise NotImplementedError("MelSpectrogramTransform: power must be 1 or 2")
        if self.win_length > n_fft:
            raise ValueError("win_length must be <= n_fft")
        window = window_fn(self.win_length) if wkwargs is None else window_fn(self.win_length, **wkwargs)
        left = (n_fft - self.win_length) // 2  # torch.stft centres a short window inside the n_fft frame
        padded = torch.zeros(n_fft)
        padded[left: left + self.win_length] = window.float()
        self.register_buffer("window", padded, persistent=False)
        self.register_buffer("fb", melscale_fbanks(n_fft // 2 + 1, self.f_min, self.f_max, n_mels, sample_rate, norm,
                                                   mel_scale).contiguous(), persistent=False)
        self.log_floor = 1e-5  # transforms/mel.py:25

    def _run(self, wav, want_spec, want_mel):
        _abi.require_cuda(wav, "MelSpectrogramTransform")
        lead = wav.shape[:-1]
        x = wav.reshape(-1, wav.shape[-1]).float().contiguous()
        B, L = x.shape
        frames = 1 + L // self.hop_length
        dev = x.device
        if self.window.device != dev:
            self.to(dev)
        spec = torch.empty(B, self.n_fft // 2 + 1, frames, device=dev) if want_spec else None
        mel = torch.empty(B, self.n_mels, frames, device=dev) if want_mel else None
        with torch.cuda.device(dev):
            _abi.check(_abi.lib().pttspp_mel_spectrogram(
                _abi.ptr(x), B, L, self.n_fft, self.hop_length, _abi.ptr(self.window), _abi.ptr(self.fb), self.n_mels,
                int(self.power), self.log_floor, None if spec is None else _abi.ptr(spec),
                None if mel is None else _abi.ptr(mel), _abi.stream_ptr(dev)))
        if spec is not None:
            spec = spec.reshape(*lead, *spec.shape[1:])
        if mel is not None:
            mel = mel.reshape(*lead, *mel.shape[1:])
        return spec, mel

    def to_spec(self, wav):
        return self._run(wav, True, False)[0]

    def spec_to_mel(self, spec):
        _abi.require_cuda(spec, "MelSpectrogramTransform.spec_to_mel")
        lead = spec.shape[:-2]
        s = spec.reshape(-1, *spec.shape[-2:]).float().contiguous()
        B, n_freq, frames = s.shape
        if n_freq != self.n_fft // 2 + 1:
            raise ValueError(f"spectrogram has {n_freq} bins, expected {self.n_fft // 2 + 1}")
        if self.fb.device != s.device:
            self.to(s.device)
        mel = torch.empty(B, self.n_mels, frames, device=s.device)
        with torch.cuda.device(s.device):
            _abi.check(_abi.lib().pttspp_mel_from_spec(_abi.ptr(s), B, n_freq, frames, _abi.ptr(self.fb), self.n_mels,
                                                       self.log_floor, _abi.ptr(mel), _abi.stream_ptr(s.device)))
        return mel.reshape(*lead, self.n_mels, frames)

    def to_mel(self, wav):
        return self._run(wav, False, True)[1]

    def forward(self, wav):
        return self.to_mel(wav)
