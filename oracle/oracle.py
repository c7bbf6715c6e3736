"""CPU oracle: a functional fp32 restatement of the PromptTTS++ inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under promptttspp_b200/ imports this file; it is used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
checker and the timed CPU baseline, never as the product path.

It restates, in plain torch-CPU functional ops on a reference-format state_dict, what the
reference's nn.Modules compute in inference; every function cites the reference file:line it
follows (paths relative to the reference repo, line/promptttspp @ a78fe65).  Parity is PINNED:
tests/test_oracle_golden.py checks this file against golden vectors produced by running the
reference's own modules (tests/golden/make_golden.py, executed where /root/reference exists).

Layouts follow the reference: activations [B, C, T] unless stated otherwise.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------


def _conv_weight(sd, prefix):
    """Plain `weight` or the weight-norm pair (torch.nn.utils.weight_norm, dim=0)."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    return torch._weight_norm(sd[prefix + ".weight_v"], sd[prefix + ".weight_g"], 0)


def sequence_mask(length, max_length=None):
    """promptttspp/utils/model.py:30-34"""
    if max_length is None:
        max_length = int(length.max())
    return torch.arange(int(max_length), device=length.device).unsqueeze(0) < length.unsqueeze(1)


# --------------------------------------------------------------------------------------------
# BigVGAN (promptttspp/vocoders/bigvgan.py, promptttspp/layers/activations.py)
# --------------------------------------------------------------------------------------------


def aa_activation(x, log_alpha, up_f, down_f):
    """AntiAliasActivation.forward, as written (activations.py:29-33, 88-96, 41-44, 115-119).

    x [B, C, L]; log_alpha [1, C, 1]; up_f / down_f [1, 1, 12].
    """
    C = x.shape[1]
    k = up_f.shape[-1]
    ratio = 2
    pad = k // ratio - 1
    pad_left = pad * ratio + (k - ratio) // 2
    pad_right = pad * ratio + (k - ratio + 1) // 2
    u = F.pad(x, (pad, pad), mode="replicate")
    u = ratio * F.conv_transpose1d(u, up_f.expand(C, -1, -1), stride=ratio, groups=C)
    u = u[..., pad_left:-pad_right]
    alpha = log_alpha.exp()
    s = u + (1.0 / (alpha + 1e-9)) * (u * alpha).sin().pow(2)
    kd = down_f.shape[-1]
    s = F.pad(s, (kd // 2 - int(kd % 2 == 0), kd // 2), mode="replicate")
    return F.conv1d(s, down_f.expand(C, -1, -1), stride=ratio, groups=C)


def aa_activation_closed_form(x, log_alpha, up_f, down_f):
    """The closed form the CUDA kernel implements (csrc/aa_snake.cu), written with gathers:
        u[2j]   = 2 sum_d x[clamp(j-3+d)] f[11-2d],  u[2j+1] = 2 sum_d x[clamp(j-2+d)] f[10-2d]
        y[t]    = sum_k s[clamp(2t+k-5, 0, 2L-1)] g[k]
    Used by the tests to pin the derivation against `aa_activation`."""
    B, C, L = x.shape
    f = up_f.reshape(-1)
    g = down_f.reshape(-1)
    j = torch.arange(L)
    ue = torch.zeros_like(x)
    uo = torch.zeros_like(x)
    for d in range(6):
        ue = ue + x[..., (j - 3 + d).clamp(0, L - 1)] * f[11 - 2 * d]
        uo = uo + x[..., (j - 2 + d).clamp(0, L - 1)] * f[10 - 2 * d]
    u = torch.stack([2 * ue, 2 * uo], dim=-1).reshape(B, C, 2 * L)
    alpha = log_alpha.exp()
    s = u + (1.0 / (alpha + 1e-9)) * (u * alpha).sin().pow(2)
    t = torch.arange(L)
    y = torch.zeros_like(x)
    for k in range(12):
        y = y + s[..., (2 * t + k - 5).clamp(0, 2 * L - 1)] * g[k]
    return y


def _aa(sd, prefix, x):
    return aa_activation(x, sd[prefix + ".act.alpha"], sd[prefix + ".up.filter"], sd[prefix + ".down.lowpass.filter"])


def bigvgan_forward(sd, cfg, mel):
    """BigVGAN.forward (bigvgan.py:120-131) with AMPLayer.forward (:42-47).  mel [B, 80, T]."""
    rates, ksz = cfg["upsample_rates"], cfg["upsample_kernel_sizes"]
    rks, rds = cfg["resblock_kernel_sizes"], cfg["resblock_dilations"]
    x = F.conv1d(mel, _conv_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(rates, ksz)):
        p = f"upsamples.{i}"
        x = F.conv_transpose1d(x, _conv_weight(sd, p), sd[p + ".bias"], stride=u, padding=u // 2 + u % 2,
                               output_padding=u % 2)
        xs = 0
        for j, (rk, rd) in enumerate(zip(rks, rds)):
            y = x
            for l, d in enumerate(rd):
                lp = f"mrfs.{i}.{j}.layers.{l}"
                h = _aa(sd, lp + ".act1", y)
                h = F.conv1d(h, _conv_weight(sd, lp + ".conv1"), sd[lp + ".conv1.bias"],
                             padding=(rk * d - d) // 2, dilation=d)
                h = _aa(sd, lp + ".act2", h)
                h = F.conv1d(h, _conv_weight(sd, lp + ".conv2"), sd[lp + ".conv2.bias"], padding=rk // 2)
                y = y + h
            xs = xs + y
        x = xs / len(rks)
    x = _aa(sd, "act_post", x)
    x = F.conv1d(x, _conv_weight(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def lowpass_filter(x, fs=100, cutoff=20, N=5):
    """utils/model.py:164-196, the torch branch (app.py:77 applies it to log-f0): scipy.signal.butter coefficients,
    short-input pass-through, torchaudio.functional.filtfilt(x, a, b, clamp=False)."""
    from scipy import signal
    from torchaudio.functional import filtfilt

    nyquist = fs // 2
    b, a = signal.butter(N, [cutoff / nyquist], "lowpass")
    if x.shape[-1] <= max(len(a), len(b)) * (N // 2 + 1):
        return x
    return filtfilt(x, torch.from_numpy(a).float(), torch.from_numpy(b).float(), clamp=False)


MEL_CFG = dict(sample_rate=24000, n_fft=512, win_length=480, hop_length=240, power=1, f_min=63, f_max=12000, n_mels=80,
               mel_scale="slaney", norm="slaney", center=True)  # egs/proposed/bin/conf/transforms/mel.yaml


def mel_filterbank(n_freqs, f_min, f_max, n_mels, sample_rate):
    """Slaney-scale, slaney-normalised triangular filters [n_freqs, n_mels] -- torchaudio.functional.melscale_fbanks
    (third-party: torchaudio, unpinned in the reference's setup.py; the MelScale buffer of transforms/mel.py:23-24)."""
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0

    def hz_to_mel(f):
        return min_log_hz / f_sp + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2)
    f_pts = f_sp * m_pts
    log_t = m_pts >= min_log_hz / f_sp
    f_pts[log_t] = min_log_hz * torch.exp(logstep * (m_pts[log_t] - min_log_hz / f_sp))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.clamp(torch.min(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]), min=0.0)
    return fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)


def mel_spectrogram(wav, cfg=None):
    """transforms/mel.py:18-34: wav [B, L] -> (|STFT| [B, n_fft/2+1, frames], log-mel [B, n_mels, frames]).  Explicit
    framing (reflect padding of n_fft/2, periodic Hann of win_length centred in the n_fft frame, hop) + rfft per frame,
    magnitude, filterbank contraction, clamp_min(1e-5).log()."""
    cfg = cfg or MEL_CFG
    n_fft, win, hop = cfg["n_fft"], cfg["win_length"], cfg["hop_length"]
    x = F.pad(wav.unsqueeze(1), (n_fft // 2, n_fft // 2), mode="reflect").squeeze(1)
    frames = x.unfold(-1, n_fft, hop)                                    # [B, frames, n_fft]
    window = torch.zeros(n_fft)
    left = (n_fft - win) // 2
    window[left:left + win] = torch.hann_window(win)
    spec = torch.fft.rfft(frames * window, dim=-1).abs()
    if cfg["power"] == 2:
        spec = spec ** 2
    spec = spec.transpose(1, 2)                                          # [B, n_freq, frames]
    fb = mel_filterbank(n_fft // 2 + 1, cfg["f_min"], cfg["f_max"], cfg["n_mels"], cfg["sample_rate"])
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)
    return spec, mel.clamp_min(1e-5).log()


def nsf_source(sd, f0_up, rand_ini, noise, sampling_rate=24000.0, harmonic_num=8, sine_amp=0.1, noise_std=0.003,
               voiced_threshold=0.0):
    """SourceModuleHnNSF.forward (vocoders/nsf.py:193-206) over SineGen.forward (:116-148) and SineGen._f02sine
    (:55-85, the non-pulse branch), with the reference's random draws injected: rand_ini [B, H] (:64-67, column 0
    zeroed), noise [B, L, H] (:143).  f0_up: [B, L, 1] -> har_source [B, L, 1]."""
    H = harmonic_num + 1
    f0_buf = torch.zeros(f0_up.shape[0], f0_up.shape[1], H)
    f0_buf[:, :, 0] = f0_up[:, :, 0]
    for idx in range(harmonic_num):
        f0_buf[:, :, idx + 1] = f0_buf[:, :, 0] * (idx + 2)
    rad = (f0_buf / sampling_rate) % 1
    ini = rand_ini.clone()
    ini[:, 0] = 0
    rad[:, 0, :] = rad[:, 0, :] + ini
    over = torch.cumsum(rad, 1) % 1
    wrap = (over[:, 1:, :] - over[:, :-1, :]) < 0
    shift = torch.zeros_like(rad)
    shift[:, 1:, :] = wrap * -1.0
    sines = torch.sin(torch.cumsum(rad + shift, dim=1) * 2 * np.pi)
    sine_waves = sines * sine_amp
    uv = torch.ones_like(f0_up) * (f0_up > voiced_threshold)
    noise_amp = uv * noise_std + (1 - uv) * sine_amp / 3
    sine_waves = sine_waves * uv + noise_amp * noise
    return torch.tanh(F.linear(sine_waves, sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"]))


def bigvgan_f0_forward(sd, cfg, mel, f0, rand_ini, noise, sampling_rate=24000.0, harmonic_num=8):
    """F0AwareBigVGAN.forward (vocoders/bigvgan_f0.py:98-115).  mel [B, 80, T], f0 [B, 1, T]; nn.Upsample(nearest,
    scale_factor=prod(rates)) of f0 (:99), the source (:100-101), then per stage x = up(x) + noise_conv(har) (:104-106)."""
    rates, ksz = cfg["upsample_rates"], cfg["upsample_kernel_sizes"]
    rks, rds = cfg["resblock_kernel_sizes"], cfg["resblock_dilations"]
    hop = int(np.prod(rates))
    f0_up = F.interpolate(f0, scale_factor=float(hop)).transpose(-1, -2)
    har = nsf_source(sd, f0_up, rand_ini, noise, sampling_rate, harmonic_num).transpose(-1, -2)
    x = F.conv1d(mel, _conv_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(rates, ksz)):
        p = f"upsamples.{i}"
        x = F.conv_transpose1d(x, _conv_weight(sd, p), sd[p + ".bias"], stride=u, padding=u // 2 + u % 2,
                               output_padding=u % 2)
        if i + 1 < len(rates):
            sf = int(np.prod(rates[i + 1:]))
            x = x + F.conv1d(har, sd[f"noise_convs.{i}.weight"], sd[f"noise_convs.{i}.bias"], stride=sf, padding=sf // 2)
        else:
            x = x + F.conv1d(har, sd[f"noise_convs.{i}.weight"], sd[f"noise_convs.{i}.bias"])
        xs = 0
        for j, (rk, rd) in enumerate(zip(rks, rds)):
            y = x
            for l, d in enumerate(rd):
                lp = f"mrfs.{i}.{j}.layers.{l}"
                h = _aa(sd, lp + ".act1", y)
                h = F.conv1d(h, _conv_weight(sd, lp + ".conv1"), sd[lp + ".conv1.bias"],
                             padding=(rk * d - d) // 2, dilation=d)
                h = _aa(sd, lp + ".act2", h)
                h = F.conv1d(h, _conv_weight(sd, lp + ".conv2"), sd[lp + ".conv2.bias"], padding=rk // 2)
                y = y + h
            xs = xs + y
        x = xs / len(rks)
    x = _aa(sd, "act_post", x)
    x = F.conv1d(x, _conv_weight(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------------------------
# Conformer text encoder (promptttspp/modules/esp/**)
# --------------------------------------------------------------------------------------------


def sinusoid(positions, d_model):
    """sin/cos(position * div_term) in fp32 (esp/transformer/embedding.py:68-77)."""
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    ang = positions.to(torch.float32).unsqueeze(1) * div_term
    pe = torch.zeros(positions.numel(), d_model)
    pe[:, 0::2] = torch.sin(ang)
    pe[:, 1::2] = torch.cos(ang)
    return pe


def rel_pos_table(T, d_model, legacy, max_len=5000):
    """legacy: reversed max_len table sliced from the front -> pe[k] = sinusoid(max_len-1-k), k < T
    (embedding.py:58-79, 234-257).  new: row k <-> relative position T-1-k (embedding.py:283-331)."""
    if legacy:
        n = max(max_len, T)
        return sinusoid(torch.arange(n - 1, -1, -1.0)[:T], d_model)
    pos = sinusoid(torch.arange(0, T, dtype=torch.float32), d_model)
    neg = sinusoid(-1 * torch.arange(0, T, dtype=torch.float32), d_model)
    return torch.cat([torch.flip(pos, [0]), neg[1:]], dim=0)


def rel_shift_legacy(bd):
    """Closed form of LegacyRelPositionMultiHeadedAttention.rel_shift (attention.py:142-162):
    out[i,j] = bd[i, T-1-i+j] (j<=i); 0 (j==i+1); bd[i+1, j-i-2] (j>=i+2)."""
    T = bd.shape[-1]
    i = torch.arange(T).unsqueeze(1)
    j = torch.arange(T).unsqueeze(0)
    low = j <= i
    row = torch.where(low, i, (i + 1).clamp(max=T - 1)).expand(T, T)
    col = torch.where(low, T - 1 - i + j, (j - i - 2).clamp(min=0))
    out = bd[..., row, col]
    return out.masked_fill((j == i + 1), 0.0)


def rel_shift_new(bd):
    """Closed form of RelPositionMultiHeadedAttention.rel_shift (attention.py:237-260):
    out[i,j] = bd[i, T-1+j-i], bd [.., T, 2T-1]."""
    T = bd.shape[-2]
    i = torch.arange(T).unsqueeze(1)
    j = torch.arange(T).unsqueeze(0)
    return bd[..., i.expand(T, T), T - 1 + j - i]


def relpos_attention(sd, p, x, pos_emb, mask2d, heads, legacy):
    """(Legacy)RelPositionMultiHeadedAttention.forward + forward_attention
    (attention.py:164-206 / 262-305, 63-93).  x [B, T, C]; pos_emb [Tp, C]; mask2d [B, T, T] bool."""
    B, T, C = x.shape
    dk = C // heads
    q = F.linear(x, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, T, heads, dk)
    k = F.linear(x, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(B, T, heads, dk).transpose(1, 2)
    v = F.linear(x, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(B, T, heads, dk).transpose(1, 2)
    pp = F.linear(pos_emb, sd[p + "linear_pos.weight"]).view(1, -1, heads, dk).transpose(1, 2)
    qu = (q + sd[p + "pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[p + "pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    bd = torch.matmul(qv, pp.transpose(-2, -1))
    bd = rel_shift_legacy(bd) if legacy else rel_shift_new(bd)
    scores = (ac + bd) / math.sqrt(dk)
    m = mask2d.unsqueeze(1).eq(0)
    scores = scores.masked_fill(m, torch.finfo(scores.dtype).min)
    attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
    o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(o, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def _ff(sd, p, x, mask):
    """MultiLayeredConv1d.forward (multi_layer_conv.py:52-67). x [B, T, C], mask [B, T, 1]."""
    k = sd[p + "w_1.weight"].shape[-1]
    x = x * mask
    h = torch.relu(F.conv1d(x.transpose(-1, 1), sd[p + "w_1.weight"], sd[p + "w_1.bias"], padding=(k - 1) // 2))
    h = h.transpose(-1, 1) * mask
    o = F.conv1d(h.transpose(-1, 1), sd[p + "w_2.weight"], sd[p + "w_2.bias"], padding=(k - 1) // 2)
    return o.transpose(-1, 1) * mask


def _conv_module(sd, p, x, mask):
    """ConvolutionModule.forward (conformer/convolution.py:58-85); Swish (swish.py:18)."""
    C = x.shape[-1]
    x = x.transpose(1, 2)
    m = mask.transpose(1, 2)
    x = F.conv1d(x, sd[p + "pointwise_conv1.weight"], sd[p + "pointwise_conv1.bias"]) * m
    x = F.glu(x, dim=1)
    k = sd[p + "depthwise_conv.weight"].shape[-1]
    x = F.conv1d(x, sd[p + "depthwise_conv.weight"], sd[p + "depthwise_conv.bias"], padding=(k - 1) // 2, groups=C) * m
    x = F.batch_norm(x, sd[p + "norm.running_mean"], sd[p + "norm.running_var"], sd[p + "norm.weight"],
                     sd[p + "norm.bias"], training=False, eps=1e-5)
    x = x * torch.sigmoid(x)
    x = F.conv1d(x, sd[p + "pointwise_conv2.weight"], sd[p + "pointwise_conv2.bias"]) * m
    return x.transpose(1, 2)


def _ln(sd, p, x, eps=1e-12):
    C = x.shape[-1]
    return F.layer_norm(x, (C,), sd[p + ".weight"], sd[p + ".bias"], eps)


def conformer_encoder(sd, cfg, emb, lens):
    """ConformerEncoder.forward (esp/__init__.py:47-65) -> Encoder.forward (conformer/encoder.py:248-282)
    -> EncoderLayer.forward (conformer/encoder_layer.py:74-162).  emb [B, T, C] -> [B, T, C]."""
    B, T, C = emb.shape
    legacy = cfg["rel_pos_type"] in (None, "legacy")
    pad = sequence_mask(lens, T)
    mask2d = pad.unsqueeze(-2) & pad.unsqueeze(-1)
    mask_ = mask2d[:, 0:1, :].transpose(1, 2).to(emb.dtype)
    x = emb * math.sqrt(C)
    pos_emb = rel_pos_table(T, C, legacy)
    for i in range(cfg["num_blocks"]):
        p = f"encoder.encoder.encoders.{i}."
        x = x * mask_
        x = x + 0.5 * _ff(sd, p + "feed_forward_macaron.", _ln(sd, p + "norm_ff_macaron", x), mask_)
        att = relpos_attention(sd, p + "self_attn.", _ln(sd, p + "norm_mha", x), pos_emb, mask2d,
                               cfg["attention_heads"], legacy)
        x = x + att * mask_
        x = x + _conv_module(sd, p + "conv_module.", _ln(sd, p + "norm_conv", x), mask_) * mask_
        x = x + 0.5 * _ff(sd, p + "feed_forward.", _ln(sd, p + "norm_ff", x), mask_) * mask_
        x = _ln(sd, p + "norm_final", x) * mask_
    x = _ln(sd, "encoder.encoder.after_norm", x)
    return x * mask2d[:, :, 0:1].to(x.dtype)


# --------------------------------------------------------------------------------------------
# variance adaptor, MDN, frame prior (promptttspp/modules/{variance_adaptor,mdn,frame_prior}.py)
# --------------------------------------------------------------------------------------------


def channel_layernorm(x, gamma, beta, eps=1e-5):
    """layers/norm.py:26-32.  x [B, C, T]; gamma/beta [1, C, 1]."""
    mean = torch.mean(x, dim=1, keepdim=True)
    var = torch.mean((x - mean) ** 2, dim=1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def _predictor_stack(sd, p, x, mask, n_layers):
    """PredictorLayer x n (variance_adaptor.py:23-36)."""
    for i in range(n_layers):
        lp = f"{p}layers.{i}."
        k = sd[lp + "conv.weight"].shape[-1]
        x = torch.relu(F.conv1d(x, sd[lp + "conv.weight"], sd[lp + "conv.bias"], padding=k // 2))
        x = channel_layernorm(x, sd[lp + "norm.gamma"], sd[lp + "norm.beta"]) * mask
    return x


def mdn_forward(sd, p, x, G, D):
    """MDNLayer.forward, dim_wise (mdn.py:50-78).  x [B, T, Cin] -> 3 x [B, T, G, D]."""
    B = x.shape[0]
    log_pi = F.log_softmax(F.linear(x, sd[p + "log_pi.weight"], sd[p + "log_pi.bias"]).view(B, -1, G, D), dim=2)
    log_sigma = F.linear(x, sd[p + "log_sigma.weight"], sd[p + "log_sigma.bias"]).view(B, -1, G, D)
    mu = F.linear(x, sd[p + "mu.weight"], sd[p + "mu.bias"]).view(B, -1, G, D)
    return log_pi, log_sigma, mu


def mdn_most_probable(log_pi, log_sigma, mu):
    """mdn_get_most_probable_sigma_and_mu, dim_wise (mdn.py:178-223): first arg-max over G."""
    idx = torch.max(log_pi, dim=2)[1].unsqueeze(2)
    return torch.exp(torch.gather(log_sigma, 2, idx).squeeze(2)), torch.gather(mu, 2, idx).squeeze(2)


def duration_log(sd, cfg, x, phone_mask):
    """MDNPredictor.infer (variance_adaptor.py:83-102). x [B, C, Tx] -> log_d [B, 1, Tx]."""
    p = "variance_adaptor.duration_predictor."
    h = _predictor_stack(sd, p, x, phone_mask, cfg["dur_layers"])
    out = mdn_forward(sd, p + "out_layer.", h.transpose(-1, -2), cfg["dur_gaussians"], 1)
    sigma, mu = mdn_most_probable(*out)
    return (mu + sigma.pow(2).clamp_min(1e-14) / 2).transpose(-1, -2)


def quantize_durations(log_d, phone_mask):
    """variance_adaptor.py:179-183: exp -> round (half to even) -> clamp_min(1) -> long -> * mask."""
    d = log_d.exp().round().clamp_min(1).long() * phone_mask.long()
    return d, d.squeeze(1).sum(dim=-1)


def length_regulate_dense(x, duration, phone_mask, frame_mask):
    """generate_path + bmm, as written (utils/model.py:37-47, variance_adaptor.py:184-187)."""
    path_mask = (phone_mask.unsqueeze(-1) * frame_mask.unsqueeze(2)).squeeze(1)
    b, t_x, t_y = path_mask.shape
    cum = torch.cumsum(duration.squeeze(1), dim=1)
    path = sequence_mask(cum.view(b * t_x), t_y).to(path_mask.dtype).view(b, t_x, t_y)
    path = path - F.pad(path, [0, 0, 1, 0, 0, 0])[:, :-1]
    return x @ (path * path_mask)


def length_regulate_indices(duration, t_y):
    """Closed form: idx[b, t] = searchsorted(cumsum(d[b]), t, right=True), -1 past the end."""
    cum = torch.cumsum(duration.reshape(duration.shape[0], -1), dim=1)
    t = torch.arange(t_y).unsqueeze(0).expand(cum.shape[0], -1).contiguous()
    idx = torch.searchsorted(cum, t, right=True)
    return torch.where(t < cum[:, -1:], idx, torch.full_like(idx, -1))


def frame_prior(sd, cfg, x, frame_mask):
    """FramePriorNetwork.forward (frame_prior.py:79-92) + PositionalEncoding (modules/embedding.py:80-92)."""
    p = "variance_adaptor.frame_prior_network."
    C, T = x.shape[1], x.shape[2]
    x = x * frame_mask
    pe = sinusoid(torch.arange(0, max(T, 1), dtype=torch.float32), C)[:T]
    x = (x.transpose(1, 2) * math.sqrt(C) + pe.unsqueeze(0)).transpose(1, 2)

    def ln(t, name):
        t = F.layer_norm(t.transpose(1, -1), (C,), sd[p + name + ".gamma"], sd[p + name + ".beta"], 1e-5)
        return t.transpose(1, -1)

    x = ln(x, "norm_emb")
    for i in range(cfg["fp_layers"]):
        k = sd[p + f"convs.{i}.weight"].shape[-1]
        res = F.gelu(F.conv1d(x * frame_mask, sd[p + f"convs.{i}.weight"], sd[p + f"convs.{i}.bias"], padding=k // 2))
        x = ln(x + res, f"norms.{i}")
    return x * frame_mask


def pitch_predict(sd, cfg, x, frame_mask):
    """Predictor.forward (variance_adaptor.py:50-59)."""
    p = "variance_adaptor.pitch_predictor."
    h = _predictor_stack(sd, p, x, frame_mask, cfg["pitch_layers"])
    out = F.conv1d(h, sd[p + "out_layer.weight"], sd[p + "out_layer.bias"]) * frame_mask
    return out.split(1, dim=1)


# --------------------------------------------------------------------------------------------
# diffusion decoder (promptttspp/modules/{diffusion,denoiser}.py)
# --------------------------------------------------------------------------------------------


def diffnet(sd, cfg, x, t, cond):
    """DiffNet.forward + ResidualBlock.forward (denoiser.py:121-143, 69-83, 34-41, 23-25).
    x [B, 80, T]; t [B] long; cond [B, C, T]."""
    p = "decoder.denoise_fn."
    C = cfg["diff_channels"]
    L = cfg["diff_layers"]
    x = torch.relu(F.conv1d(x, sd[p + "input_projection.weight"], sd[p + "input_projection.bias"]))
    half = C // 2
    emb = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    emb = cfg.get("diff_scale", 1) * t[:, None] * emb[None, :]
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    h = F.linear(emb, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])
    h = h * torch.tanh(F.softplus(h))
    step = F.linear(h, sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])
    skips = []
    for l in range(L):
        lp = f"{p}residual_layers.{l}."
        dil = 2 ** (l % cfg["diff_dilation_cycle"])
        k = sd[lp + "dilated_conv.weight"].shape[-1]
        d = F.linear(step, sd[lp + "diffusion_projection.weight"], sd[lp + "diffusion_projection.bias"]).unsqueeze(-1)
        c = F.conv1d(cond, sd[lp + "conditioner_projection.weight"], sd[lp + "conditioner_projection.bias"])
        y = F.conv1d(x + d, sd[lp + "dilated_conv.weight"], sd[lp + "dilated_conv.bias"],
                     padding=(k * dil - dil) // 2, dilation=dil) + c
        gate, filt = torch.chunk(y, 2, dim=1)
        y = torch.sigmoid(gate) * torch.tanh(filt)
        y = F.conv1d(y, sd[lp + "output_projection.weight"], sd[lp + "output_projection.bias"])
        residual, skip = torch.chunk(y, 2, dim=1)
        x = (x + residual) / math.sqrt(2.0)
        skips.append(skip)
    x = torch.sum(torch.stack(skips), dim=0) / math.sqrt(L)
    x = torch.relu(F.conv1d(x, sd[p + "skip_projection.weight"], sd[p + "skip_projection.bias"]))
    return F.conv1d(x, sd[p + "output_projection.weight"], sd[p + "output_projection.bias"])


def diffusion_sample(sd, cfg, cond, x_T, z, steps=None):
    """GaussianDiffusion.inference / p_sample / p_mean_variance / q_posterior
    (diffusion.py:320-356, 210-221, 198-208, 181-196).  cond [B, C, T]; x_T [B, 80, T]; z [K, B, 80, T].
    Returns the de-normalised mel [B, T, 80] * norm_scale transposed back to [B, 80, T]."""
    K = cfg["K_step"]
    x = x_T
    B = x.shape[0]
    n = 0
    for i in reversed(range(K)):
        if steps is not None and n >= steps:
            break
        t = torch.full((B,), i, dtype=torch.long)
        eps = diffnet(sd, cfg, x, t, cond)
        x0 = sd["decoder.sqrt_recip_alphas_cumprod"][i] * x - sd["decoder.sqrt_recipm1_alphas_cumprod"][i] * eps
        x0 = x0.clamp(-1.0, 1.0)
        mean = sd["decoder.posterior_mean_coef1"][i] * x0 + sd["decoder.posterior_mean_coef2"][i] * x
        nonzero = 0.0 if i == 0 else 1.0
        x = mean + nonzero * (0.5 * sd["decoder.posterior_log_variance_clipped"][i]).exp() * z[n]
        n += 1
    if cfg.get("norm_scale"):
        return x * cfg["norm_scale"]
    return (x + 1) / 2 * (cfg["a_max"] - cfg["a_min"]) + cfg["a_min"]


# --------------------------------------------------------------------------------------------
# whole acoustic model (promptttspp/models/prompttts_mdn_v2_final/model.py)
# --------------------------------------------------------------------------------------------

ACOUSTIC_CFG = dict(
    rel_pos_type="legacy", num_blocks=4, attention_heads=2, dur_layers=2, dur_gaussians=4, pitch_layers=5,
    fp_layers=6, style_gaussians=10, norm_style_emb=True, K_step=100, diff_layers=20, diff_channels=256,
    diff_dilation_cycle=4, diff_scale=1, norm_scale=6.0, a_min=0, a_max=20,
)

VOCODER_CFG = dict(
    upsample_rates=[6, 5, 4, 2], upsample_kernel_sizes=[12, 10, 8, 4], resblock_kernel_sizes=[3, 7, 11],
    resblock_dilations=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
)


def style_embedding(sd, cfg, cls_emb, z_style, noise_scale, comp_u=None):
    """PromptEncoder adaptor (prompt_encoder.py:45-56) + normalise + style MDN + sample_style_emb
    (model.py:284-296, 185-196).  cls_emb [B, 768]; z_style [B, 1, C] -> [B, C, 1]."""
    p = "prompt_encoder.adaptor."
    e = torch.relu(F.linear(cls_emb, sd[p + "0.weight"], sd[p + "0.bias"]))
    e = torch.relu(F.linear(e, sd[p + "2.weight"], sd[p + "2.bias"]))
    e = F.linear(e, sd[p + "4.weight"], sd[p + "4.bias"]).unsqueeze(-1)
    if cfg["norm_style_emb"]:
        e = F.normalize(e, dim=1)
    C = e.shape[1]
    log_pi, log_sigma, mu = mdn_forward(sd, "style_mdn.", e.transpose(-1, -2), cfg["style_gaussians"], C)
    if comp_u is None:
        sigma, mu = mdn_most_probable(log_pi, log_sigma, mu)
    else:
        # mdn_sample_sigma_and_mu (mdn.py:226-257), dim-wise: Categorical(probs=pi).sample() per (b, dimension), with
        # the draw realised as the inverse CDF of the injected uniforms comp_u [B, C]
        pi = log_pi.exp().squeeze(1).transpose(1, 2)                      # (B, C, G)
        cdf = torch.cumsum(pi / pi.sum(-1, keepdim=True), dim=-1)
        idx = (comp_u.unsqueeze(-1) >= cdf).sum(-1).clamp(max=pi.shape[-1] - 1)  # (B, C)
        one_hot = F.one_hot(idx, pi.shape[-1]).unsqueeze(1).transpose(2, 3)
        mu_s = torch.sum(mu * one_hot, dim=2)
        sigma = torch.exp(torch.sum(log_sigma * one_hot, dim=2))
        mu = mu_s
    style = mu + sigma * z_style * noise_scale
    if cfg["norm_style_emb"]:
        style = F.normalize(style, dim=-1)
    return style.transpose(-1, -2)


def bert_forward(sd, input_ids, attention_mask, heads, eps=1e-12, prefix=""):
    """HF transformers BertModel(...).last_hidden_state (the reference's third-party dependency, prompt_encoder.py:25,
    33-38; transformers 5.5.0 modeling_bert.py): BertEmbeddings (word + token_type(0) + position, LayerNorm), then per
    layer BertSelfAttention (eager: softmax(q k^T / sqrt(dk) + (1 - mask) * finfo.min) v), BertSelfOutput (dense +
    residual, LayerNorm), BertIntermediate (dense, exact-erf GELU), BertOutput (dense + residual, LayerNorm)."""
    e = prefix + "embeddings."
    B, T = input_ids.shape
    x = F.embedding(input_ids, sd[e + "word_embeddings.weight"]) + sd[e + "token_type_embeddings.weight"][0]
    x = x + sd[e + "position_embeddings.weight"][:T]
    Hd = x.shape[-1]
    x = F.layer_norm(x, (Hd,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps)
    dk = Hd // heads
    add = torch.zeros(B, 1, 1, T)
    if attention_mask is not None:
        add = (1.0 - attention_mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    n = 0
    while f"{prefix}encoder.layer.{n}.attention.self.query.weight" in sd:
        p = f"{prefix}encoder.layer.{n}."
        lin = lambda t, name: F.linear(t, sd[p + name + ".weight"], sd[p + name + ".bias"])
        q = lin(x, "attention.self.query").view(B, T, heads, dk).transpose(1, 2)
        k = lin(x, "attention.self.key").view(B, T, heads, dk).transpose(1, 2)
        v = lin(x, "attention.self.value").view(B, T, heads, dk).transpose(1, 2)
        w = F.softmax(q @ k.transpose(-1, -2) / math.sqrt(dk) + add, dim=-1)
        ctx = (w @ v).transpose(1, 2).reshape(B, T, Hd)
        x = F.layer_norm(lin(ctx, "attention.output.dense") + x, (Hd,), sd[p + "attention.output.LayerNorm.weight"],
                         sd[p + "attention.output.LayerNorm.bias"], eps)
        h = F.gelu(lin(x, "intermediate.dense"))
        x = F.layer_norm(lin(h, "output.dense") + x, (Hd,), sd[p + "output.LayerNorm.weight"],
                         sd[p + "output.LayerNorm.bias"], eps)
        n += 1
    return x


def style_encoder(sd, speech, in_lens=None, prefix="reference_encoder.", conv_layers=6, stride=2, heads=4):
    """StyleEncoder.forward (modules/style_encoder.py:70-80): ReferenceEncoder (reference_encoder.py:95-124: 6 x
    [Conv2d k3 s2 no bias, BatchNorm2d eval, ReLU], GRU whose last VALID state is the embedding -- the packed-sequence
    branch :112-121) and StyleTokenLayer + MultiHeadedAttention (style_encoder.py:106-171).
    speech [B, 80, Lmax], in_lens [B] -> [B, 256, 1]."""
    p = prefix + "ref_enc."
    x = speech.transpose(1, 2).unsqueeze(1)
    for i in range(conv_layers):
        cp, bp = f"{p}convs.{3 * i}.", f"{p}convs.{3 * i + 1}."
        x = F.conv2d(x, sd[cp + "weight"], None, stride=stride, padding=1)
        x = F.batch_norm(x, sd[bp + "running_mean"], sd[bp + "running_var"], sd[bp + "weight"], sd[bp + "bias"],
                         training=False, eps=1e-5)
        x = torch.relu(x)
    hs = x.transpose(1, 2)
    B, T = hs.shape[0], hs.shape[1]
    hs = hs.contiguous().view(B, T, -1)
    lens = torch.full((B,), T, dtype=torch.long)
    if in_lens is not None:
        lens = torch.clamp(torch.ceil(in_lens.float() / (stride ** conv_layers)).long(), 1)
    w_ih, w_hh = sd[p + "gru.weight_ih_l0"], sd[p + "gru.weight_hh_l0"]
    b_ih, b_hh = sd[p + "gru.bias_ih_l0"], sd[p + "gru.bias_hh_l0"]
    Hn = w_hh.shape[1]
    ref = torch.zeros(B, Hn)
    for b in range(B):  # torch.nn.GRU cell equations, gate order (r, z, n)
        h = torch.zeros(Hn)
        for t in range(int(lens[b])):
            gi = F.linear(hs[b, t], w_ih, b_ih)
            gh = F.linear(h, w_hh, b_hh)
            r = torch.sigmoid(gi[:Hn] + gh[:Hn])
            zg = torch.sigmoid(gi[Hn:2 * Hn] + gh[Hn:2 * Hn])
            n = torch.tanh(gi[2 * Hn:] + r * gh[2 * Hn:])
            h = (1 - zg) * n + zg * h
        ref[b] = h
    q = prefix + "stl."
    gst = torch.tanh(sd[q + "gst_embs"]).unsqueeze(0).expand(B, -1, -1)
    m = q + "mha."
    Fd = sd[m + "linear_q.weight"].shape[0]
    dk = Fd // heads
    qv = F.linear(ref.unsqueeze(1), sd[m + "linear_q.weight"], sd[m + "linear_q.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    kv = F.linear(gst, sd[m + "linear_k.weight"], sd[m + "linear_k.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    vv = F.linear(gst, sd[m + "linear_v.weight"], sd[m + "linear_v.bias"]).view(B, -1, heads, dk).transpose(1, 2)
    score = F.softmax((qv @ kv.transpose(-1, -2)) / math.sqrt(dk * heads), dim=-1)
    o = (score @ vv).transpose(-1, -2).contiguous().view(B, 1, -1)
    o = F.linear(o, sd[m + "linear_out.weight"], sd[m + "linear_out.bias"])
    return o.squeeze(1).unsqueeze(-1)


def acoustic_infer_batch(sd, cfg, phoneme, phone_lengths, cls_emb, z_style, x_T=None, z=None, noise_scale=1.0,
                         noise_fn=None, steps=None, return_intermediates=False, reference_mel=None, ref_lengths=None,
                         comp_u=None):
    """PromptTTSMDNDurCFG.infer_batch (model.py:261-325) with use_max=True and injected noise.

    x_T / z may be None, then `noise_fn(shape)` is called in the reference's order once Ty is known.
    Returns mel [B, 80, Ty], log_cf0, vuv [B, 1, Ty], frame_lengths (float) [+ dict of intermediates]."""
    phone_mask = sequence_mask(phone_lengths).unsqueeze(1).to(phoneme.dtype)
    x = F.embedding(phoneme, sd["phoneme_emb.emb.weight"], padding_idx=0)
    if cfg.get("emb_do_scale"):
        x = x * math.sqrt(x.shape[-1])
    x = x.transpose(-1, -2) * phone_mask
    x = conformer_encoder(sd, cfg, x.transpose(1, 2), phone_lengths).transpose(1, 2)
    if reference_mel is not None:  # model.py:296-301
        style = style_encoder(sd, reference_mel, ref_lengths)
        if cfg["norm_style_emb"]:
            style = F.normalize(style, dim=1)
    else:
        style = style_embedding(sd, cfg, cls_emb, z_style, noise_scale, comp_u)
    x = x + style
    enc_state = x
    pm = phone_mask.to(x.dtype)
    log_d = duration_log(sd, cfg, x, pm)
    duration, frame_lengths = quantize_durations(log_d, phone_mask)
    frame_mask = sequence_mask(frame_lengths).unsqueeze(1).to(x.dtype)
    x = length_regulate_dense(x, duration.to(x.dtype), pm, frame_mask)
    x = frame_prior(sd, cfg, x, frame_mask)
    log_cf0, vuv = pitch_predict(sd, cfg, x, frame_mask)
    pe = "variance_adaptor.pitch_emb."
    x = x + F.conv1d(log_cf0, sd[pe + "weight"], sd[pe + "bias"]) * frame_mask
    B, Ty = x.shape[0], x.shape[2]
    if x_T is None:
        x_T = noise_fn((B, 80, Ty))
        z = torch.stack([noise_fn((B, 80, Ty)) for _ in range(cfg["K_step"] if steps is None else steps)])
    mel = diffusion_sample(sd, cfg, x, x_T, z, steps=steps) * frame_mask
    out = (mel, log_cf0, vuv, frame_mask.sum(dim=(1, 2)))
    if return_intermediates:
        return out + (dict(enc_state=enc_state, log_d=log_d, duration=duration, cond=x, style=style),)
    return out
