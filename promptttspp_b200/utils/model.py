"""Host-side helpers that callers of the reference import next to the models.

Reference: promptttspp/utils/model.py:23-34, 164-196.
"""
import numpy as np
import torch
from torch import nn


def remove_weight_norm_(m):
    """``module.apply(remove_weight_norm_)`` as egs/proposed/bin/synthesize.py:108,116 does.

    Tolerated on every module of this package: modules without weight-norm are
    left untouched; the packed kernel weights are rebuilt lazily afterwards.
    """
    try:
        nn.utils.remove_weight_norm(m)
    except ValueError:
        return


def sequence_mask(length, max_length=None):
    if max_length is None:
        max_length = length.max()
    pos = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return pos.unsqueeze(0) < length.unsqueeze(1)


def butter_lowpass(N, Wn):
    """Digital Butterworth low-pass (b, a) like scipy.signal.butter(N, [Wn], "lowpass") (utils/model.py:183): analog
    prototype poles on the unit circle, pre-warped cutoff, bilinear transform.  Host side, float64."""
    k = np.arange(1, N + 1)
    poles = np.exp(1j * np.pi * (2 * k + N - 1) / (2 * N))          # analog prototype, cutoff 1 rad/s
    warped = 4.0 * np.tan(np.pi * Wn / 2.0)                          # fs = 2 in scipy's convention
    p = poles * warped
    gain = warped ** N
    pz = (4.0 + p) / (4.0 - p)                                       # bilinear transform, fs = 2
    kz = gain * np.real(1.0 / np.prod(4.0 - p))
    b = kz * np.poly(-np.ones(N))
    a = np.real(np.poly(pz))
    return b, a


def lowpass_filter(x, fs=100, cutoff=20, N=5, lengths=None):
    """Zero-phase low-pass of a CUDA tensor along its last axis -- promptttspp/utils/model.py:164-196 (app.py:77
    smooths log-f0 with it before the vocoder).  Same short-input pass-through; the filtering itself is
    pttspp_iir_filtfilt (torchaudio.functional.filtfilt semantics).

    `lengths` (not in the reference, which is only ever called per utterance): valid samples per row of a zero-padded
    batch.  Each row is then filtered exactly as a call on x[r, ..., :lengths[r]] would be -- without it the backward
    pass of a short row would start inside the padding and smear the step at the row's end over its last frames."""
    from .. import _abi

    nyquist = fs // 2
    b, a = butter_lowpass(N, cutoff / nyquist)
    x_len = x.shape[-1]
    min_len = max(len(a), len(b)) * (N // 2 + 1)
    if x_len <= min_len:
        return x
    if not isinstance(x, torch.Tensor):
        raise NotImplementedError("lowpass_filter: only torch CUDA tensors are supported (no host fallback)")
    _abi.require_cuda(x, "lowpass_filter")
    xf = x.contiguous().float()
    rows = xf.numel() // x_len
    y = torch.empty_like(xf)
    tmp = torch.empty_like(xf)
    bd = torch.tensor(b, dtype=torch.float32, device=x.device)
    ad = torch.tensor(a, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        if lengths is None:
            _abi.check(_abi.lib().pttspp_iir_filtfilt(_abi.ptr(xf), _abi.ptr(y), _abi.ptr(tmp), rows, x_len, _abi.ptr(bd),
                                                      _abi.ptr(ad), len(b), _abi.stream_ptr(x.device)))
        else:
            lens = lengths.to(device=x.device, dtype=torch.int64).reshape(-1)
            if lens.numel() != x.shape[0]:
                raise ValueError(f"{lens.numel()} lengths for {x.shape[0]} batch rows")
            lens = lens.repeat_interleave(rows // x.shape[0]).contiguous()  # one entry per filtered row
            _abi.check(_abi.lib().pttspp_iir_filtfilt_ragged(
                _abi.ptr(xf), _abi.ptr(y), _abi.ptr(tmp), rows, x_len, _abi.ptr(lens), min_len, _abi.ptr(bd),
                _abi.ptr(ad), len(b), _abi.stream_ptr(x.device)))
    return y.view(x.shape)
