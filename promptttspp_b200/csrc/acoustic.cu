// PromptTTSMDNDurCFG inference (promptttspp/models/prompttts_mdn_v2_final/model.py:261-325) as two
// stream-ordered launch sequences over channels-last activations:
//   encode : embedding -> Conformer encoder -> prompt adaptor + style MDN sample -> x + style ->
//            MDN duration predictor -> integer durations / frame lengths
//   decode : length regulator -> frame prior -> pitch predictor -> + pitch embedding ->
//            K_step x DiffNet (20 gated residual layers) + DDPM posterior update -> mel
// Everything that depends only on the weights is folded at finalize(): BatchNorm(eval) affine,
// the diffusion-step embedding MLP (a [K_step][layers][C] table), q/k/v weight concatenation, the
// concatenated conditioner projections of all residual layers (step-invariant, hoisted out of the
// sampling loop), gate/filter channel interleaving.
#include "common.h"
#include "diffnet_layer.h"

namespace pttspp {

struct LNW {
  float* gamma = nullptr;
  float* beta = nullptr;
};

struct EncBlockW {
  LNW norm_ff_macaron, norm_mha, norm_conv, norm_ff, norm_final;
  PackedConv mac_w1, mac_w2, ff_w1, ff_w2;
  PackedConv qkv, pos, out;
  float* bias_u = nullptr;
  float* bias_v = nullptr;
  PackedConv pw1, pw2;
  float* dw_w = nullptr;
  float* dw_b = nullptr;
  float* bn_scale = nullptr;
  float* bn_shift = nullptr;
};

struct PredLayerW {
  PackedConv conv;
  LNW norm;
};

struct DiffLayerW {
  PackedConv dilated;  // C -> 2C, gate/filter interleaved
  PackedConv outp;     // C -> 2C: columns [0,C) residual, [C,2C) skip
};

// bump allocator over the caller's workspace; with base == nullptr it only measures
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base((uint8_t*)b) {}
  template <typename T>
  T* take(int64_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + off) : nullptr;
    off += (size_t)std::max<int64_t>(n, 1) * sizeof(T);
    return p;
  }
};

}  // namespace pttspp

struct pttspp_acoustic {
  pttspp_acoustic_config cfg;
  pttspp::TensorStore store;
  pttspp::DeviceBuffers dev;
  bool finalized = false;
  float* emb = nullptr;
  std::vector<pttspp::EncBlockW> blocks;
  pttspp::LNW after_norm;
  pttspp::PackedConv ad0, ad1, ad2, mdn_pi, mdn_ls, mdn_mu;
  std::vector<pttspp::PredLayerW> dur_layers, pitch_layers;
  float *dur_wpi = nullptr, *dur_bpi = nullptr, *dur_wls = nullptr, *dur_bls = nullptr, *dur_wmu = nullptr,
        *dur_bmu = nullptr;
  float *pitch_out_w = nullptr, *pitch_out_b = nullptr, *pitch_emb_w = nullptr, *pitch_emb_b = nullptr;
  pttspp::LNW fp_norm_emb;
  std::vector<pttspp::PredLayerW> fp_layers;
  // diffusion
  pttspp::PackedConv in_proj, in_proj_tc, cond_all, skip_proj, out_proj;
  std::vector<pttspp::DiffLayerW> diff;
  float* step_table = nullptr;  // [K_step][layers][C]
  bool use_umma = true;         // tcgen05 split-fp16 path for the DiffNet contractions (PTTSPP_DISABLE_UMMA=1: off)
  // text side (encoder feed-forward k9 convs) on the chunked near-fp32 tcgen05 path; PTTSPP_TEXT_TC=0: fp32 CUDA cores
  bool text_tc = true;
  // all residual layers of one diffusion step in ONE persistent kernel (csrc/diffnet_layer.cu); PTTSPP_DIFFNET_FUSED=0
  // keeps the two-launches-per-layer path (the only one for other channel counts / kernel sizes)
  pttspp::DiffNetStack diffnet;
  bool use_fused = false;
  // the step boundary (skip projection -> output projection -> DDPM update -> next input projection) as one kernel
  // (csrc/diffnet_tail.cu); PTTSPP_DIFFNET_TAIL=0 keeps the four separate launches
  pttspp::DiffNetTail tail;
  bool use_tail = false;
  bool h_fp32 = false;          // PTTSPP_H_FP32=1: fp32 copy of the residual stream (two-launch path, A/B measurements)
  std::vector<float> c_recip, c_recipm1, coef1, coef2, logvar;
};

namespace pttspp {
namespace {

LNW load_ln(const TensorStore& st, DeviceBuffers& dev, const std::string& gname, const std::string& bname, int C) {
  LNW w;
  w.gamma = dev.upload(st.get(gname, C).data);
  w.beta = dev.upload(st.get(bname, C).data);
  return w;
}

void run_ln(const LNW& w, const float* in, const float* in2, float* out, int B, int T, int C, float eps,
            const int64_t* in_len, const int64_t* out_len, float in_scale, const float* row_add, cudaStream_t s) {
  pttspp_layernorm_desc d;
  memset(&d, 0, sizeof(d));
  d.in = in; d.in2 = in2; d.row_add = row_add; d.gamma = w.gamma; d.beta = w.beta; d.out = out;
  d.bs = (int64_t)T * C; d.ld = C; d.B = B; d.T = T; d.C = C; d.eps = eps; d.in_scale = in_scale;
  d.in_len = in_len; d.out_len = out_len;
  layernorm_cl(d, s);
}

// concatenate several [Cout_i][Cin](xK) torch weights along Cout into one packed conv
PackedConv load_concat(const TensorStore& st, DeviceBuffers& dev, const std::vector<std::string>& prefixes, int Cout_each,
                       int Cin, int K, bool interleave_each, bool has_bias) {
  const int n = (int)prefixes.size();
  PackedConv c;
  c.Cin = Cin; c.Cout = Cout_each * n; c.K = K; c.dil = 1; c.pad = (K - 1) / 2;
  c.w_ld = round_up(c.Cout, 4);
  std::vector<float> packed((size_t)K * Cin * c.w_ld, 0.f), bias((size_t)c.Cout, 0.f);
  const int wl = round_up(Cout_each, 4);
  std::vector<float> tmp((size_t)K * Cin * wl);
  const int half = Cout_each / 2;
  for (int i = 0; i < n; ++i) {
    const HostTensor& w = st.get(prefixes[i] + ".weight", (int64_t)Cout_each * Cin * K);
    pack_conv_weight(w.data.data(), nullptr, Cout_each, Cin, K, tmp.data(), wl, interleave_each, 0);
    for (size_t r = 0; r < (size_t)K * Cin; ++r)
      std::copy(tmp.begin() + r * wl, tmp.begin() + r * wl + Cout_each, packed.begin() + r * c.w_ld + (size_t)i * Cout_each);
    if (has_bias) {
      const HostTensor& b = st.get(prefixes[i] + ".bias", Cout_each);
      for (int co = 0; co < Cout_each; ++co) {
        const int col = interleave_each ? (co < half ? 2 * co : 2 * (co - half) + 1) : co;
        bias[(size_t)i * Cout_each + col] = b.data[co];
      }
    }
  }
  c.w = dev.upload(packed);
  if (has_bias) c.bias = dev.upload(bias);
  if (K == 1 && Cin % 64 == 0) {
    // split-fp16 planes [Cout][Cin] of the same concatenation for the tcgen05 path (rows already in packed column order)
    std::vector<float> rows((size_t)c.Cout * Cin);
    for (int co = 0; co < c.Cout; ++co)
      for (int ci = 0; ci < Cin; ++ci) rows[(size_t)co * Cin + ci] = packed[(size_t)ci * c.w_ld + co];
    std::vector<uint16_t> hi(rows.size()), lo(rows.size());
    pack_conv_weight_split(rows.data(), nullptr, c.Cout, Cin, 1, hi.data(), lo.data(), 0, &c.w_scale_inv);
    c.w_hi = dev.upload_bytes(hi.data(), hi.size() * 2);
    c.w_lo = dev.upload_bytes(lo.data(), lo.size() * 2);
  }
  return c;
}

// Diffusion-step embedding (denoiser.py:34-41, 103-105, 132-133, 70): depends only on the integer
// step, so the sinusoid -> Linear -> Mish -> Linear -> per-layer Linear chain becomes a table.
std::vector<float> build_step_table(const TensorStore& st, const pttspp_acoustic_config& c) {
  const int C = c.diff_channels, H = 4 * C, K = c.K_step, L = c.diff_layers;
  const std::string p = "decoder.denoise_fn.";
  const HostTensor& w0 = st.get(p + "mlp.0.weight", (int64_t)H * C);
  const HostTensor& b0 = st.get(p + "mlp.0.bias", H);
  const HostTensor& w2 = st.get(p + "mlp.2.weight", (int64_t)C * H);
  const HostTensor& b2 = st.get(p + "mlp.2.bias", C);
  std::vector<float> table((size_t)K * L * C);
  const int half = C / 2;
  const float emb_c = logf(10000.f) / (float)(half - 1);
  std::vector<float> e(C), h1(H), h2(C);
  for (int t = 0; t < K; ++t) {
    for (int i = 0; i < half; ++i) {
      const float f = expf((float)i * -emb_c);
      const float arg = c.diff_scale * (float)t * f;
      e[i] = sinf(arg);
      e[half + i] = cosf(arg);
    }
    for (int o = 0; o < H; ++o) {
      double acc = b0.data[o];
      const float* wr = w0.data.data() + (size_t)o * C;
      for (int i = 0; i < C; ++i) acc += (double)wr[i] * e[i];
      const float x = (float)acc;
      const float sp = x > 20.f ? x : log1pf(expf(x));  // F.softplus (threshold 20)
      h1[o] = x * tanhf(sp);                            // Mish (denoiser.py:23-25)
    }
    for (int o = 0; o < C; ++o) {
      double acc = b2.data[o];
      const float* wr = w2.data.data() + (size_t)o * H;
      for (int i = 0; i < H; ++i) acc += (double)wr[i] * h1[i];
      h2[o] = (float)acc;
    }
    for (int l = 0; l < L; ++l) {
      const std::string lp = p + "residual_layers." + std::to_string(l) + ".diffusion_projection";
      const HostTensor& w = st.get(lp + ".weight", (int64_t)C * C);
      const HostTensor& b = st.get(lp + ".bias", C);
      float* dst = table.data() + ((size_t)t * L + l) * C;
      for (int o = 0; o < C; ++o) {
        double acc = b.data[o];
        const float* wr = w.data.data() + (size_t)o * C;
        for (int i = 0; i < C; ++i) acc += (double)wr[i] * h2[i];
        dst[o] = (float)acc;
      }
    }
  }
  return table;
}

struct EncodeWs {
  float *x, *y, *hff, *qkv, *att, *pw, *cm, *p, *bd, *e1, *e2, *e3, *lp, *ls, *mu, *style, *logd;
  uint16_t *yh, *yl, *fh, *fl;  // operand planes of the feed-forward input [B][Tx][C] and hidden layer [B][Tx][units]
};

EncodeWs carve_encode(const pttspp_acoustic_config& c, int B, int Tx, int Tp, Carver& cv) {
  EncodeWs w;
  const int C = c.channels;
  const int64_t n = (int64_t)B * Tx;
  w.x = cv.take<float>(n * C);
  w.y = cv.take<float>(n * C);
  w.hff = cv.take<float>(n * c.enc_linear_units);
  w.qkv = cv.take<float>(n * 3 * C);
  w.att = cv.take<float>(n * C);
  w.pw = cv.take<float>(n * 2 * C);
  w.cm = cv.take<float>(n * C);
  w.p = cv.take<float>((int64_t)Tp * C);
  w.bd = cv.take<float>((int64_t)B * c.enc_heads * Tx * Tp);
  w.e1 = cv.take<float>((int64_t)B * c.prompt_mid);
  w.e2 = cv.take<float>((int64_t)B * c.prompt_mid);
  w.e3 = cv.take<float>((int64_t)B * C);
  w.lp = cv.take<float>((int64_t)B * c.style_gaussians * C);
  w.ls = cv.take<float>((int64_t)B * c.style_gaussians * C);
  w.mu = cv.take<float>((int64_t)B * c.style_gaussians * C);
  w.style = cv.take<float>((int64_t)B * C);
  w.logd = cv.take<float>(n);
  w.yh = cv.take<uint16_t>(n * C);
  w.yl = cv.take<uint16_t>(n * C);
  w.fh = cv.take<uint16_t>(n * c.enc_linear_units);
  w.fl = cv.take<uint16_t>(n * c.enc_linear_units);
  return w;
}

struct DecodeWs {
  float *xa, *xb, *tmp, *lcf0, *vuv, *condp, *xt, *h, *z, *skip, *s, *eps;
  uint16_t *yh, *yl, *zh, *zl, *sh, *sl, *ph, *pl;  // split-fp16 operand planes [B][Ty][DC]
  uint16_t *xh, *xl;                                // x_t planes [B][Ty][round_up(mel, 64)]
  uint16_t *yh2, *yl2;                              // second pair of residual-stream planes (fused stack: ping-pong)
  unsigned* done;                                   // fused stack: per (layer, unit) completion counters
  float* znoise;                                    // [B][mel][Ty] noise of the current step (decode_rng)
};

DecodeWs carve_decode(const pttspp_acoustic_config& c, int B, int Ty, Carver& cv, size_t flags_bytes = 0) {
  DecodeWs w;
  const int C = c.channels, DC = c.diff_channels;
  const int64_t n = (int64_t)B * Ty;
  w.xa = cv.take<float>(n * C);
  w.xb = cv.take<float>(n * C);
  w.tmp = cv.take<float>(n * C);
  w.lcf0 = cv.take<float>(n);
  w.vuv = cv.take<float>(n);
  w.condp = cv.take<float>(n * 2 * DC * c.diff_layers);
  w.xt = cv.take<float>(n * c.mel_dim);
  w.h = cv.take<float>(n * DC);
  w.z = cv.take<float>(n * DC);
  w.skip = cv.take<float>(n * DC);
  w.s = cv.take<float>(n * DC);
  w.eps = cv.take<float>(n * c.mel_dim);
  w.yh = cv.take<uint16_t>(n * DC); w.yl = cv.take<uint16_t>(n * DC);
  w.zh = cv.take<uint16_t>(n * DC); w.zl = cv.take<uint16_t>(n * DC);
  w.sh = cv.take<uint16_t>(n * DC); w.sl = cv.take<uint16_t>(n * DC);
  w.ph = cv.take<uint16_t>(n * DC); w.pl = cv.take<uint16_t>(n * DC);
  w.xh = cv.take<uint16_t>(n * round_up(c.mel_dim, 64)); w.xl = cv.take<uint16_t>(n * round_up(c.mel_dim, 64));
  w.yh2 = cv.take<uint16_t>(flags_bytes ? n * DC : 0); w.yl2 = cv.take<uint16_t>(flags_bytes ? n * DC : 0);
  w.done = cv.take<unsigned>((int64_t)(flags_bytes / sizeof(unsigned)));
  w.znoise = cv.take<float>(n * c.mel_dim);
  return w;
}

}  // namespace
}  // namespace pttspp

using namespace pttspp;

extern "C" int pttspp_acoustic_create(const pttspp_acoustic_config* cfg, pttspp_acoustic_t** out) {
  PT_API_BEGIN
  PT_CHECK(cfg && out, "null argument");
  PT_CHECK(cfg->channels % 32 == 0 && cfg->channels <= 1024, "acoustic: channels=%d unsupported", cfg->channels);
  PT_CHECK(cfg->enc_heads >= 1 && cfg->channels % cfg->enc_heads == 0 && (cfg->channels / cfg->enc_heads) % 32 == 0,
           "acoustic: attention head size must be a multiple of 32");
  PT_CHECK(cfg->enc_ff_kernel % 2 == 1 && cfg->enc_cnn_kernel % 2 == 1 && cfg->dur_kernel % 2 == 1 &&
               cfg->pitch_kernel % 2 == 1 && cfg->fp_kernel % 2 == 1 && cfg->diff_kernel % 2 == 1,
           "acoustic: kernel sizes must be odd");
  PT_CHECK(cfg->mel_dim % 16 == 0 && cfg->prompt_in % 16 == 0 && cfg->prompt_mid % 16 == 0 &&
               cfg->enc_linear_units % 16 == 0 && cfg->diff_channels % 16 == 0,
           "acoustic: channel counts must be multiples of 16");
  PT_CHECK(cfg->diff_channels == cfg->channels, "acoustic: residual_channels must equal encoder_hidden_dim");
  PT_CHECK(cfg->K_step >= 1 && cfg->diff_layers >= 1 && cfg->diff_dilation_cycle >= 1, "acoustic: bad diffusion config");
  auto* h = new pttspp_acoustic();
  h->cfg = *cfg;
  *out = h;
  PT_API_END
}

extern "C" void pttspp_acoustic_destroy(pttspp_acoustic_t* h) { delete h; }

extern "C" int pttspp_acoustic_set_tensor(pttspp_acoustic_t* h, const char* name, const float* data,
                                          const int64_t* shape, int ndim, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(h, "null handle");
  h->store.set(name, data, shape, ndim, (cudaStream_t)stream);
  h->finalized = false;
  PT_API_END
}

extern "C" int pttspp_acoustic_finalize(pttspp_acoustic_t* h, pttspp_stream_t) {
  PT_API_BEGIN
  PT_CHECK(h, "null handle");
  const auto& c = h->cfg;
  const TensorStore& st = h->store;
  DeviceBuffers& dev = h->dev;
  dev.release();
  h->blocks.clear(); h->dur_layers.clear(); h->pitch_layers.clear(); h->fp_layers.clear(); h->diff.clear();
  const int C = c.channels;

  h->emb = dev.upload(st.get("phoneme_emb.emb.weight", (int64_t)c.num_vocab * C).data);
  for (int i = 0; i < c.enc_blocks; ++i) {
    const std::string p = "encoder.encoder.encoders." + std::to_string(i) + ".";
    EncBlockW b;
    b.norm_ff_macaron = load_ln(st, dev, p + "norm_ff_macaron.weight", p + "norm_ff_macaron.bias", C);
    b.norm_mha = load_ln(st, dev, p + "norm_mha.weight", p + "norm_mha.bias", C);
    b.norm_conv = load_ln(st, dev, p + "norm_conv.weight", p + "norm_conv.bias", C);
    b.norm_ff = load_ln(st, dev, p + "norm_ff.weight", p + "norm_ff.bias", C);
    b.norm_final = load_ln(st, dev, p + "norm_final.weight", p + "norm_final.bias", C);
    const int kf = c.enc_ff_kernel, U = c.enc_linear_units;
    b.mac_w1 = load_conv1d(st, dev, p + "feed_forward_macaron.w_1", U, C, kf, 1, (kf - 1) / 2);
    b.mac_w2 = load_conv1d(st, dev, p + "feed_forward_macaron.w_2", C, U, kf, 1, (kf - 1) / 2);
    b.ff_w1 = load_conv1d(st, dev, p + "feed_forward.w_1", U, C, kf, 1, (kf - 1) / 2);
    b.ff_w2 = load_conv1d(st, dev, p + "feed_forward.w_2", C, U, kf, 1, (kf - 1) / 2);
    if (C % 64 == 0 && U % 64 == 0) {  // split-fp16 planes for the chunked (near-fp32) tcgen05 path of the text side
      attach_split_weights(st, dev, p + "feed_forward_macaron.w_1", b.mac_w1, false);
      attach_split_weights(st, dev, p + "feed_forward_macaron.w_2", b.mac_w2, false);
      attach_split_weights(st, dev, p + "feed_forward.w_1", b.ff_w1, false);
      attach_split_weights(st, dev, p + "feed_forward.w_2", b.ff_w2, false);
    }
    b.qkv = load_concat(st, dev, {p + "self_attn.linear_q", p + "self_attn.linear_k", p + "self_attn.linear_v"}, C, C, 1,
                        false, true);
    b.pos = load_linear(st, dev, p + "self_attn.linear_pos", C, C, false);
    b.out = load_linear(st, dev, p + "self_attn.linear_out", C, C, true);
    b.bias_u = dev.upload(st.get(p + "self_attn.pos_bias_u", C).data);
    b.bias_v = dev.upload(st.get(p + "self_attn.pos_bias_v", C).data);
    b.pw1 = load_conv1d(st, dev, p + "conv_module.pointwise_conv1", 2 * C, C, 1, 1, 0);
    b.pw2 = load_conv1d(st, dev, p + "conv_module.pointwise_conv2", C, C, 1, 1, 0);
    b.dw_w = dev.upload(st.get(p + "conv_module.depthwise_conv.weight", (int64_t)C * c.enc_cnn_kernel).data);
    b.dw_b = dev.upload(st.get(p + "conv_module.depthwise_conv.bias", C).data);
    {  // eval-mode BatchNorm1d as y*scale + shift (eps 1e-5), the form torch's inference path uses
      const auto& g = st.get(p + "conv_module.norm.weight", C).data;
      const auto& be = st.get(p + "conv_module.norm.bias", C).data;
      const auto& rm = st.get(p + "conv_module.norm.running_mean", C).data;
      const auto& rv = st.get(p + "conv_module.norm.running_var", C).data;
      std::vector<float> sc(C), sh(C);
      for (int k = 0; k < C; ++k) {
        const float invstd = 1.f / sqrtf(rv[k] + 1e-5f);
        sc[k] = g[k] * invstd;
        sh[k] = be[k] - rm[k] * sc[k];
      }
      b.bn_scale = dev.upload(sc);
      b.bn_shift = dev.upload(sh);
    }
    h->blocks.push_back(b);
  }
  h->after_norm = load_ln(st, dev, "encoder.encoder.after_norm.weight", "encoder.encoder.after_norm.bias", C);

  h->ad0 = load_linear(st, dev, "prompt_encoder.adaptor.0", c.prompt_mid, c.prompt_in);
  h->ad1 = load_linear(st, dev, "prompt_encoder.adaptor.2", c.prompt_mid, c.prompt_mid);
  h->ad2 = load_linear(st, dev, "prompt_encoder.adaptor.4", C, c.prompt_mid);
  const int GD = c.style_gaussians * C;
  h->mdn_pi = load_linear(st, dev, "style_mdn.log_pi", GD, C);
  h->mdn_ls = load_linear(st, dev, "style_mdn.log_sigma", GD, C);
  h->mdn_mu = load_linear(st, dev, "style_mdn.mu", GD, C);

  const std::string va = "variance_adaptor.";
  for (int i = 0; i < c.dur_layers; ++i) {
    const std::string p = va + "duration_predictor.layers." + std::to_string(i) + ".";
    PredLayerW l;
    l.conv = load_conv1d(st, dev, p + "conv", C, C, c.dur_kernel, 1, c.dur_kernel / 2);
    l.norm = load_ln(st, dev, p + "norm.gamma", p + "norm.beta", C);
    h->dur_layers.push_back(l);
  }
  {
    const std::string p = va + "duration_predictor.out_layer.";
    const int G = c.dur_gaussians;
    h->dur_wpi = dev.upload(st.get(p + "log_pi.weight", (int64_t)G * C).data);
    h->dur_bpi = dev.upload(st.get(p + "log_pi.bias", G).data);
    h->dur_wls = dev.upload(st.get(p + "log_sigma.weight", (int64_t)G * C).data);
    h->dur_bls = dev.upload(st.get(p + "log_sigma.bias", G).data);
    h->dur_wmu = dev.upload(st.get(p + "mu.weight", (int64_t)G * C).data);
    h->dur_bmu = dev.upload(st.get(p + "mu.bias", G).data);
  }
  for (int i = 0; i < c.pitch_layers; ++i) {
    const std::string p = va + "pitch_predictor.layers." + std::to_string(i) + ".";
    PredLayerW l;
    l.conv = load_conv1d(st, dev, p + "conv", C, C, c.pitch_kernel, 1, c.pitch_kernel / 2);
    if (C % 64 == 0) attach_split_weights(st, dev, p + "conv", l.conv, false);
    l.norm = load_ln(st, dev, p + "norm.gamma", p + "norm.beta", C);
    h->pitch_layers.push_back(l);
  }
  h->pitch_out_w = dev.upload(st.get(va + "pitch_predictor.out_layer.weight", 2 * C).data);
  h->pitch_out_b = dev.upload(st.get(va + "pitch_predictor.out_layer.bias", 2).data);
  h->pitch_emb_w = dev.upload(st.get(va + "pitch_emb.weight", C).data);
  h->pitch_emb_b = dev.upload(st.get(va + "pitch_emb.bias", C).data);
  h->fp_norm_emb = load_ln(st, dev, va + "frame_prior_network.norm_emb.gamma", va + "frame_prior_network.norm_emb.beta", C);
  for (int i = 0; i < c.fp_layers; ++i) {
    PredLayerW l;
    l.conv = load_conv1d(st, dev, va + "frame_prior_network.convs." + std::to_string(i), C, C, c.fp_kernel, 1,
                         c.fp_kernel / 2);
    if (C % 64 == 0) attach_split_weights(st, dev, va + "frame_prior_network.convs." + std::to_string(i), l.conv, false);
    const std::string n = va + "frame_prior_network.norms." + std::to_string(i);
    l.norm = load_ln(st, dev, n + ".gamma", n + ".beta", C);
    h->fp_layers.push_back(l);
  }

  // ---- diffusion decoder ----
  const std::string dn = "decoder.denoise_fn.";
  const int DC = c.diff_channels;
  h->in_proj = load_conv1d(st, dev, dn + "input_projection", DC, c.mel_dim, 1, 1, 0);
  {
    // tensor-core copy with the mel axis zero-padded to a multiple of 64 (80 -> 128): x_t travels as padded planes
    const int Mp = round_up(c.mel_dim, 64);
    const auto& wt = st.get(dn + "input_projection.weight", (int64_t)DC * c.mel_dim).data;  // [DC][mel][1]
    std::vector<float> wp((size_t)DC * Mp, 0.f);
    for (int co = 0; co < DC; ++co)
      for (int ci = 0; ci < c.mel_dim; ++ci) wp[(size_t)co * Mp + ci] = wt[(size_t)co * c.mel_dim + ci];
    std::vector<uint16_t> hi(wp.size()), lo(wp.size());
    pack_conv_weight_split(wp.data(), nullptr, DC, Mp, 1, hi.data(), lo.data(), 0, &h->in_proj_tc.w_scale_inv);
    h->in_proj_tc = h->in_proj;
    h->in_proj_tc.Cin = Mp;
    pack_conv_weight_split(wp.data(), nullptr, DC, Mp, 1, hi.data(), lo.data(), 0, &h->in_proj_tc.w_scale_inv);
    h->in_proj_tc.w_hi = dev.upload_bytes(hi.data(), hi.size() * 2);
    h->in_proj_tc.w_lo = dev.upload_bytes(lo.data(), lo.size() * 2);
  }
  std::vector<std::string> cond_names;
  for (int l = 0; l < c.diff_layers; ++l) {
    const std::string p = dn + "residual_layers." + std::to_string(l) + ".";
    DiffLayerW w;
    const int dl = 1 << (l % c.diff_dilation_cycle);
    w.dilated = load_conv1d(st, dev, p + "dilated_conv", 2 * DC, DC, c.diff_kernel, dl,
                            (c.diff_kernel * dl - dl) / 2, /*interleave=*/true);
    w.outp = load_conv1d(st, dev, p + "output_projection", 2 * DC, DC, 1, 1, 0);
    attach_split_weights(st, dev, p + "dilated_conv", w.dilated, /*interleave=*/true);
    attach_split_weights(st, dev, p + "output_projection", w.outp, false);
    h->diff.push_back(w);
    cond_names.push_back(p + "conditioner_projection");
  }
  h->cond_all = load_concat(st, dev, cond_names, 2 * DC, C, 1, /*interleave_each=*/true, true);
  h->skip_proj = load_conv1d(st, dev, dn + "skip_projection", DC, DC, 1, 1, 0);
  h->out_proj = load_conv1d(st, dev, dn + "output_projection", c.mel_dim, DC, 1, 1, 0);
  attach_split_weights(st, dev, dn + "skip_projection", h->skip_proj, false);
  attach_split_weights(st, dev, dn + "output_projection", h->out_proj, false);
  {
    const char* e = getenv("PTTSPP_DISABLE_UMMA");
    h->use_umma = !(e && e[0] == '1') && DC % 64 == 0;
    h->h_fp32 = getenv("PTTSPP_H_FP32") != nullptr;
    const char* tt = getenv("PTTSPP_TEXT_TC");
    h->text_tc = h->use_umma && !(tt && tt[0] == '0') && !h->blocks.empty() && h->blocks[0].mac_w1.w_hi != nullptr;
    const char* f = getenv("PTTSPP_DIFFNET_FUSED");
    const int max_dil = 1 << (std::min(c.diff_dilation_cycle, c.diff_layers) - 1);
    h->use_fused = h->use_umma && !(f && f[0] == '0') && !h->h_fp32 && h->in_proj_tc.w_hi != nullptr &&
                   DiffNetStack::supported(DC, c.diff_kernel, max_dil);
    if (h->use_fused) {
      std::vector<DiffLayerHost> hl;
      for (int l = 0; l < c.diff_layers; ++l) {
        const DiffLayerW& w = h->diff[l];
        hl.push_back(DiffLayerHost{w.dilated.w_hi, w.dilated.w_lo, w.outp.w_hi, w.outp.w_lo, w.dilated.bias, w.outp.bias,
                                   w.dilated.w_scale_inv, w.outp.w_scale_inv, w.dilated.dil});
      }
      h->diffnet.set_layers(hl);
      const char* tl = getenv("PTTSPP_DIFFNET_TAIL");
      h->use_tail = !(tl && tl[0] == '0') && DC == 256 && c.mel_dim <= 128 && c.mel_dim % 8 == 0 &&
                    round_up(c.mel_dim, 64) == 128 && h->skip_proj.w_hi && h->out_proj.w_hi;
      if (h->use_tail) {
        DiffTailHost th;
        th.wsp_hi = h->skip_proj.w_hi; th.wsp_lo = h->skip_proj.w_lo;
        th.wop_hi = h->out_proj.w_hi; th.wop_lo = h->out_proj.w_lo;
        th.wip_hi = h->in_proj_tc.w_hi; th.wip_lo = h->in_proj_tc.w_lo;
        th.bias_sp = h->skip_proj.bias; th.bias_op = h->out_proj.bias; th.bias_ip = h->in_proj_tc.bias;
        th.scale_sp = h->skip_proj.w_scale_inv / sqrtf((float)c.diff_layers);
        th.scale_op = h->out_proj.w_scale_inv;
        th.scale_ip = h->in_proj_tc.w_scale_inv;
        th.mel = c.mel_dim; th.mel_pad = round_up(c.mel_dim, 64);
        h->tail.set_weights(th);
      }
    }
  }
  h->step_table = dev.upload(build_step_table(st, c));
  h->c_recip = st.get("decoder.sqrt_recip_alphas_cumprod", c.K_step).data;
  h->c_recipm1 = st.get("decoder.sqrt_recipm1_alphas_cumprod", c.K_step).data;
  h->coef1 = st.get("decoder.posterior_mean_coef1", c.K_step).data;
  h->coef2 = st.get("decoder.posterior_mean_coef2", c.K_step).data;
  h->logvar = st.get("decoder.posterior_log_variance_clipped", c.K_step).data;
  h->finalized = true;
  PT_API_END
}

extern "C" size_t pttspp_acoustic_encode_workspace_bytes(const pttspp_acoustic_t* h, int B, int Tx) {
  if (!h || B <= 0 || Tx <= 0) return 0;
  Carver cv(nullptr);
  carve_encode(h->cfg, B, Tx, h->cfg.rel_pos_legacy ? Tx : 2 * Tx - 1, cv);
  return cv.off + 512;
}

// style_in != nullptr: the style vector comes from the reference-mel style encoder (cls_emb / z_style unused)
static void acoustic_encode_impl(pttspp_acoustic_t* h, const int64_t* phoneme, const int64_t* phone_len, int B, int Tx,
                                 const float* pos_emb, int Tp, const float* cls_emb, const float* z_style,
                                 const float* style_in, const float* comp_u, float noise_scale, int use_max,
                                 float* enc_state, int64_t* dur,
                                 int64_t* frame_len, float* log_dur, float* style_emb, void* workspace,
                                 size_t workspace_bytes, pttspp_stream_t stream) {
  {
  PT_CHECK(h && phoneme && phone_len && pos_emb && enc_state && dur && frame_len, "null argument");
  PT_CHECK(style_in || (cls_emb && z_style), "null style input");
  PT_CHECK(style_in || use_max || comp_u, "acoustic: use_max=0 needs the component draw (pttspp_acoustic_encode_sampled)");
  PT_CHECK(h->finalized, "acoustic: finalize() has not been called after the last set_tensor()");
  PT_CHECK(B >= 1 && Tx >= 1, "acoustic: empty batch (B=%d, Tx=%d)", B, Tx);
  const auto& c = h->cfg;
  PT_CHECK(Tp == (c.rel_pos_legacy ? Tx : 2 * Tx - 1), "acoustic: pos_emb has %d rows, expected %d", Tp,
           c.rel_pos_legacy ? Tx : 2 * Tx - 1);
  PT_CHECK(workspace && workspace_bytes >= pttspp_acoustic_encode_workspace_bytes(h, B, Tx), "acoustic: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  Carver cv((void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255));
  EncodeWs w = carve_encode(c, B, Tx, Tp, cv);
  const int C = c.channels, H = c.enc_heads, dk = C / H;
  const int64_t* len = phone_len;

  // phoneme embedding * mask, then the encoder front-end scale sqrt(C)  (layers/embedding.py:30-36,
  // esp/transformer/embedding.py:253)
  embedding_cl(phoneme, len, h->emb, B, Tx, C, c.num_vocab, c.emb_do_scale ? sqrtf((float)C) : 1.f, w.x, s);

  // positionwise feed-forward (multi_layer_conv.py:52-67): x += 0.5 * w_2(relu(w_1(y * m)) * m) * m with y = LN(x).
  // tcgen05 path: the masked LN output travels as split-fp16 planes, w_1 emits the masked hidden layer directly as planes,
  // both contractions accumulate in chunks of 16 tensor-core steps summed in round-to-nearest fp32 (conv1d impl 3).
  const int U = c.enc_linear_units;
  auto feed_forward = [&](const PackedConv& w1, const PackedConv& w2) {
    auto d = conv_desc(w1, w.y, B, Tx, w.hff);
    d.out_len = len; d.act = PTTSPP_ACT_RELU;
    auto e = conv_desc(w2, w.hff, B, Tx, w.x);
    e.out_len = len; e.res = w.x; e.res_bs = (int64_t)Tx * C; e.res_ld = C; e.alpha = 0.5f;
    if (h->text_tc) {
      split_f16_rows(w.y, B, Tx, C, len, w.yh, w.yl, s);
      d.in_hi = w.yh; d.in_lo = w.yl; d.in_bs = (int64_t)Tx * C; d.in_ld = C;
      d.w_hi = w1.w_hi; d.w_lo = w1.w_lo; d.w_scale_inv = w1.w_scale_inv; d.impl = 3;
      d.out = nullptr;
      d.out_hi = w.fh; d.out_lo = w.fl; d.out_plane_bs = (int64_t)Tx * U; d.out_plane_ld = U;
      e.in_hi = w.fh; e.in_lo = w.fl; e.in_bs = (int64_t)Tx * U; e.in_ld = U;
      e.w_hi = w2.w_hi; e.w_lo = w2.w_lo; e.w_scale_inv = w2.w_scale_inv; e.impl = 3;
    } else {
      d.in_len = len;
    }
    conv1d_cl(d, s);
    conv1d_cl(e, s);
  };
  for (const EncBlockW& b : h->blocks) {
    // macaron feed-forward
    run_ln(b.norm_ff_macaron, w.x, nullptr, w.y, B, Tx, C, 1e-12f, nullptr, nullptr, 1.f, nullptr, s);
    feed_forward(b.mac_w1, b.mac_w2);
    // relative-position self-attention
    run_ln(b.norm_mha, w.x, nullptr, w.y, B, Tx, C, 1e-12f, nullptr, nullptr, 1.f, nullptr, s);
    {
      auto d = conv_desc(b.qkv, w.y, B, Tx, w.qkv);
      conv1d_cl(d, s);
      auto pd = conv_desc(b.pos, pos_emb, 1, Tp, w.p);
      conv1d_cl(pd, s);
      relpos_attention(w.qkv, w.qkv + C, w.qkv + 2 * C, w.p, b.bias_u, b.bias_v, len, B, Tx, H, dk, c.rel_pos_legacy,
                       w.bd, w.att, 3 * C, s);
      auto o = conv_desc(b.out, w.att, B, Tx, w.x);
      o.out_len = len; o.res = w.x; o.res_bs = (int64_t)Tx * C; o.res_ld = C;
      conv1d_cl(o, s);
    }
    // convolution module
    run_ln(b.norm_conv, w.x, nullptr, w.y, B, Tx, C, 1e-12f, nullptr, nullptr, 1.f, nullptr, s);
    {
      auto d = conv_desc(b.pw1, w.y, B, Tx, w.pw);
      d.out_len = len;
      conv1d_cl(d, s);
      glu_dw_bn_swish_cl(w.pw, len, b.dw_w, b.dw_b, b.bn_scale, b.bn_shift, B, Tx, C, c.enc_cnn_kernel, w.cm, s);
      auto e = conv_desc(b.pw2, w.cm, B, Tx, w.x);
      e.out_len = len; e.res = w.x; e.res_bs = (int64_t)Tx * C; e.res_ld = C;
      conv1d_cl(e, s);
    }
    // feed-forward
    run_ln(b.norm_ff, w.x, nullptr, w.y, B, Tx, C, 1e-12f, nullptr, nullptr, 1.f, nullptr, s);
    feed_forward(b.ff_w1, b.ff_w2);
    run_ln(b.norm_final, w.x, nullptr, w.x, B, Tx, C, 1e-12f, nullptr, len, 1.f, nullptr, s);
  }
  run_ln(h->after_norm, w.x, nullptr, enc_state, B, Tx, C, 1e-12f, nullptr, len, 1.f, nullptr, s);

  if (style_in) {
    // reference-mel style path (model.py:232-237): normalise, broadcast-add
    float* style = style_emb ? style_emb : w.style;
    PT_CUDA(cudaMemcpyAsync(style, style_in, (size_t)B * C * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (c.norm_style_emb) l2_normalize_rows(style, B, C, s);
    add_row_broadcast(enc_state, style, B, Tx, C, s);
  } else {
    // prompt adaptor MLP on the sentence embedding, style MDN, sampled + normalised style vector
    auto d0 = conv_desc(h->ad0, cls_emb, 1, B, w.e1);
    d0.act = PTTSPP_ACT_RELU;
    conv1d_cl(d0, s);
    auto d1 = conv_desc(h->ad1, w.e1, 1, B, w.e2);
    d1.act = PTTSPP_ACT_RELU;
    conv1d_cl(d1, s);
    auto d2 = conv_desc(h->ad2, w.e2, 1, B, w.e3);
    conv1d_cl(d2, s);
    if (c.norm_style_emb) l2_normalize_rows(w.e3, B, C, s);
    auto p0 = conv_desc(h->mdn_pi, w.e3, 1, B, w.lp);
    conv1d_cl(p0, s);
    auto p1 = conv_desc(h->mdn_ls, w.e3, 1, B, w.ls);
    conv1d_cl(p1, s);
    auto p2 = conv_desc(h->mdn_mu, w.e3, 1, B, w.mu);
    conv1d_cl(p2, s);
    float* style = style_emb ? style_emb : w.style;
    style_mdn_sample(w.lp, w.ls, w.mu, z_style, B, c.style_gaussians, C, noise_scale, c.norm_style_emb, style, s,
                     use_max ? nullptr : comp_u);
    add_row_broadcast(enc_state, style, B, Tx, C, s);  // padded phoneme columns become non-zero (model.py:301)
  }

  // MDN duration predictor (variance_adaptor.py:83-102) and quantisation (:179-183)
  {
    const float* cur = enc_state;
    float* bufs[2] = {w.x, w.y};
    int bi = 0;
    for (const PredLayerW& l : h->dur_layers) {
      auto d = conv_desc(l.conv, cur, B, Tx, w.att);
      d.act = PTTSPP_ACT_RELU;
      conv1d_cl(d, s);
      run_ln(l.norm, w.att, nullptr, bufs[bi], B, Tx, C, 1e-5f, nullptr, len, 1.f, nullptr, s);
      cur = bufs[bi];
      bi ^= 1;
    }
    float* logd = log_dur ? log_dur : w.logd;
    mdn_duration_head(cur, h->dur_wpi, h->dur_bpi, h->dur_wls, h->dur_bls, h->dur_wmu, h->dur_bmu, B * Tx, C,
                      c.dur_gaussians, logd, s);
    duration_quantize(logd, len, B, Tx, dur, frame_len, s);
  }
  }
}

extern "C" int pttspp_acoustic_encode(pttspp_acoustic_t* h, const int64_t* phoneme, const int64_t* phone_len, int B,
                                      int Tx, const float* pos_emb, int Tp, const float* cls_emb, const float* z_style,
                                      float noise_scale, int use_max, float* enc_state, int64_t* dur,
                                      int64_t* frame_len, float* log_dur, float* style_emb, void* workspace,
                                      size_t workspace_bytes, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(cls_emb && z_style, "null argument");
  acoustic_encode_impl(h, phoneme, phone_len, B, Tx, pos_emb, Tp, cls_emb, z_style, nullptr, nullptr, noise_scale, use_max,
                       enc_state, dur, frame_len, log_dur, style_emb, workspace, workspace_bytes, stream);
  PT_API_END
}

extern "C" int pttspp_acoustic_encode_sampled(pttspp_acoustic_t* h, const int64_t* phoneme, const int64_t* phone_len, int B,
                                              int Tx, const float* pos_emb, int Tp, const float* cls_emb,
                                              const float* z_style, const float* comp_u, float noise_scale,
                                              float* enc_state, int64_t* dur, int64_t* frame_len, float* log_dur,
                                              float* style_emb, void* workspace, size_t workspace_bytes,
                                              pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(cls_emb && z_style && comp_u, "null argument");
  acoustic_encode_impl(h, phoneme, phone_len, B, Tx, pos_emb, Tp, cls_emb, z_style, nullptr, comp_u, noise_scale, 0,
                       enc_state, dur, frame_len, log_dur, style_emb, workspace, workspace_bytes, stream);
  PT_API_END
}

extern "C" int pttspp_acoustic_encode_ref(pttspp_acoustic_t* h, const int64_t* phoneme, const int64_t* phone_len, int B,
                                          int Tx, const float* pos_emb, int Tp, const float* style_in, float* enc_state,
                                          int64_t* dur, int64_t* frame_len, float* log_dur, void* workspace,
                                          size_t workspace_bytes, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(style_in, "null argument");
  acoustic_encode_impl(h, phoneme, phone_len, B, Tx, pos_emb, Tp, nullptr, nullptr, style_in, nullptr, 1.f, 1, enc_state, dur,
                       frame_len, log_dur, nullptr, workspace, workspace_bytes, stream);
  PT_API_END
}


extern "C" size_t pttspp_acoustic_decode_workspace_bytes(const pttspp_acoustic_t* h, int B, int Tx, int Ty) {
  (void)Tx;
  if (!h || B <= 0 || Ty <= 0) return 0;
  Carver cv(nullptr);
  carve_decode(h->cfg, B, Ty, cv, h->use_fused ? h->diffnet.flags_bytes(B, Ty) : 0);
  return cv.off + 512;
}

namespace pttspp {
uint64_t philox_normal(float* out, int64_t numel, uint64_t seed, uint64_t offset, cudaStream_t s);  // philox.cu
}

// z != nullptr: the K_step draws are given; z == nullptr: they are drawn here, step by step, from torch's Philox stream
// (rng_seed, rng_offset) -- *rng_offset_out receives the offset after the K_step draws
static int acoustic_decode_impl(pttspp_acoustic_t* h, const float* enc_state, const int64_t* dur,
                                const int64_t* frame_len, int B, int Tx, int Ty, const float* pe_abs,
                                const float* x_T, const float* z, uint64_t rng_seed, uint64_t rng_offset,
                                uint64_t* rng_offset_out, float* mel, float* log_cf0, float* vuv,
                                float* cond_out, void* workspace, size_t workspace_bytes,
                                pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(h && enc_state && dur && frame_len && pe_abs && x_T && mel, "null argument");
  PT_CHECK(h->finalized, "acoustic: finalize() has not been called after the last set_tensor()");
  PT_CHECK(B >= 1 && Tx >= 1 && Ty >= 1, "acoustic: empty batch (B=%d, Tx=%d, Ty=%d)", B, Tx, Ty);
  PT_CHECK(workspace && workspace_bytes >= pttspp_acoustic_decode_workspace_bytes(h, B, Tx, Ty),
           "acoustic: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const auto& c = h->cfg;
  Carver cv((void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255));
  const bool fused = h->use_fused;
  const size_t flags_bytes = fused ? h->diffnet.flags_bytes(B, Ty) : 0;
  DecodeWs w = carve_decode(c, B, Ty, cv, flags_bytes);
  const int C = c.channels, DC = c.diff_channels, M = c.mel_dim;
  const int64_t* flen = frame_len;
  const int64_t bsC = (int64_t)Ty * C;
  // route a conv through the tcgen05 split-fp16 path: operand planes [B][Ty][Cin] + the conv's split weight planes
  auto tc_in = [&](pttspp_conv1d_desc& q, const uint16_t* hi, const uint16_t* lo, const PackedConv& pc) {
    q.in_hi = hi; q.in_lo = lo; q.in_bs = (int64_t)Ty * pc.Cin; q.in_ld = pc.Cin;
    q.w_hi = pc.w_hi; q.w_lo = pc.w_lo; q.w_scale_inv = pc.w_scale_inv;
    q.impl = 2;
  };

  // length regulator (gather instead of the one-hot matmul of utils/model.py:37-47)
  length_regulate(enc_state, dur, B, Tx, C, Ty, w.xa, nullptr, s);

  // frame prior network (frame_prior.py:79-92): LN(x*sqrt(C) + pe), 6 x LN(x + gelu(conv(x*m))), * m
  run_ln(h->fp_norm_emb, w.xa, nullptr, w.xb, B, Ty, C, 1e-5f, nullptr, nullptr, sqrtf((float)C), pe_abs, s);
  {
    float* cur = w.xb;
    float* nxt = w.xa;
    for (size_t i = 0; i < h->fp_layers.size(); ++i) {
      const PredLayerW& l = h->fp_layers[i];
      auto d = conv_desc(l.conv, cur, B, Ty, w.tmp);
      d.act = PTTSPP_ACT_GELU;
      if (h->use_umma && l.conv.w_hi) {
        // frames are known only after the (bit-exact, fp32 CUDA-core) duration path: from here on the contractions run
        // on the tensor cores.  `x * mask` in front of the conv = zero rows in the operand planes.
        split_f16_rows(cur, B, Ty, C, flen, w.yh, w.yl, s);
        tc_in(d, w.yh, w.yl, l.conv);
      } else {
        d.in_len = flen;
      }
      conv1d_cl(d, s);
      const bool last = (i + 1 == h->fp_layers.size());
      run_ln(l.norm, cur, w.tmp, nxt, B, Ty, C, 1e-5f, nullptr, last ? flen : nullptr, 1.f, nullptr, s);
      std::swap(cur, nxt);
    }
    // cur = frame-prior output x (masked); keep it in w.xa
    if (cur != w.xa) {
      PT_CUDA(cudaMemcpyAsync(w.xa, cur, (size_t)B * bsC * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
  }
  // pitch predictor (variance_adaptor.py:50-59) -> log_cf0 / vuv, pitch embedding added to x (:192-201)
  {
    const float* cur = w.xa;
    float* bufs[2] = {w.xb, w.h};  // w.h is free until the sampling loop (same size: DC == C)
    int bi = 0;
    for (const PredLayerW& l : h->pitch_layers) {
      auto d = conv_desc(l.conv, cur, B, Ty, w.tmp);
      d.act = PTTSPP_ACT_RELU;
      if (h->use_umma && l.conv.w_hi) {
        split_f16_rows(cur, B, Ty, C, nullptr, w.yh, w.yl, s);
        tc_in(d, w.yh, w.yl, l.conv);
      }
      conv1d_cl(d, s);
      run_ln(l.norm, w.tmp, nullptr, bufs[bi], B, Ty, C, 1e-5f, nullptr, flen, 1.f, nullptr, s);
      cur = bufs[bi];
      bi ^= 1;
    }
    float* lc = log_cf0 ? log_cf0 : w.lcf0;
    float* vv = vuv ? vuv : w.vuv;
    pitch_head(cur, h->pitch_out_w, h->pitch_out_b, flen, B, Ty, C, lc, vv, s);
    pitch_embed_add(w.xa, lc, h->pitch_emb_w, h->pitch_emb_b, flen, B, Ty, C, s);
  }
  const float* cond = w.xa;
  if (cond_out) PT_CUDA(cudaMemcpyAsync(cond_out, cond, (size_t)B * bsC * sizeof(float), cudaMemcpyDeviceToDevice, s));

  // ---- DDPM ancestral sampling (diffusion.py:320-356) ----
  // conditioner projections of all residual layers: step-invariant, computed once
  const int CP = 2 * DC * c.diff_layers;
  if (fused) {
    // layer-major layout [layer][B][Ty][2*DC]: the fused kernel reads one contiguous 2 KB row per (layer, frame)
    split_f16_rows(cond, B, Ty, C, nullptr, w.yh, w.yl, s);
    for (int l = 0; l < c.diff_layers; ++l) {
      auto d = conv_desc(h->cond_all, cond, B, Ty, w.condp + (size_t)l * B * Ty * 2 * DC);
      d.Cout = 2 * DC; d.out_ld = 2 * DC; d.out_bs = (int64_t)Ty * 2 * DC;
      d.bias = h->cond_all.bias + (size_t)l * 2 * DC;
      tc_in(d, w.yh, w.yl, h->cond_all);
      d.w_hi = (const uint16_t*)h->cond_all.w_hi + (size_t)l * 2 * DC * C;
      d.w_lo = (const uint16_t*)h->cond_all.w_lo + (size_t)l * 2 * DC * C;
      conv1d_cl(d, s);
    }
    PT_CUDA(cudaMemsetAsync(w.done, 0, flags_bytes, s));
  } else {
    auto d = conv_desc(h->cond_all, cond, B, Ty, w.condp);
    if (h->use_umma && h->cond_all.w_hi) {
      split_f16_rows(cond, B, Ty, C, nullptr, w.yh, w.yl, s);
      tc_in(d, w.yh, w.yl, h->cond_all);
    }
    conv1d_cl(d, s);
  }
  transpose_bct_to_btc(x_T, w.xt, B, M, Ty, s);
  const int Mp = round_up(M, 64);
  const bool tc_inproj = h->use_umma && h->in_proj_tc.w_hi != nullptr;
  if (tc_inproj) split_f16_pad(w.xt, (int64_t)B * Ty, M, Mp, w.xh, w.xl, s);  // later steps: written by ddpm_update
  const float sqrt2 = sqrtf(2.f);
  const float inv_sqrt_layers = 1.f / sqrtf((float)c.diff_layers);
  const int64_t bsD = (int64_t)Ty * DC;
  const bool um = h->use_umma;
  // PTTSPP_H_FP32=1 keeps an fp32 copy of the residual stream (the pre-planes behaviour, for A/B measurements)
  const bool h_planes = um && tc_inproj && DC % 16 == 0 && !h->h_fp32;
  auto planes_in = [&](pttspp_conv1d_desc& q, const uint16_t* hi, const uint16_t* lo, const PackedConv& pc, int row_off) {
    q.in_hi = hi; q.in_lo = lo; q.in_bs = bsD; q.in_ld = DC;
    q.w_hi = (const uint16_t*)pc.w_hi + (size_t)row_off * pc.Cin;
    q.w_lo = (const uint16_t*)pc.w_lo + (size_t)row_off * pc.Cin;
    q.w_scale_inv = pc.w_scale_inv;
    q.impl = 2;
  };
  auto planes_out = [&](pttspp_conv1d_desc& q, uint16_t* hi, uint16_t* lo, const float* add) {
    q.out_hi = hi; q.out_lo = lo; q.out_plane_bs = bsD; q.out_plane_ld = DC; q.out_plane_add = add;
  };
  const bool tail = fused && h->use_tail && tc_inproj && h_planes;
  for (int step = c.K_step - 1; step >= 0; --step) {
    const float* step_emb = h->step_table + (size_t)step * c.diff_layers * DC;
    if (!tail || step == c.K_step - 1) {  // with the fused step boundary the previous step's tail wrote these planes
      auto d = conv_desc(tc_inproj ? h->in_proj_tc : h->in_proj, w.xt, B, Ty, w.h);
      d.act = PTTSPP_ACT_RELU;
      if (tc_inproj) tc_in(d, w.xh, w.xl, h->in_proj_tc);
      if (um) {
        planes_out(d, w.yh, w.yl, step_emb);  // y_0 = h + step_emb[0] as operand planes
        if (h_planes) d.out = nullptr;        // the residual stream travels only as those planes
      }
      conv1d_cl(d, s);
    }
    if (fused) {
      DiffNetRun r;
      r.B = B; r.T = Ty; r.layer_begin = 0; r.layer_end = c.diff_layers;
      r.cond = w.condp; r.step_emb = step_emb;
      r.y_hi[0] = w.yh; r.y_lo[0] = w.yl; r.y_hi[1] = w.yh2; r.y_lo[1] = w.yl2;
      r.skip = w.skip; r.skip_hi = w.sh; r.skip_lo = w.sl;
      r.done = w.done; r.epoch = (unsigned)(c.K_step - step); r.dbg_z = nullptr;
      const double flops = 2.0 * B * (double)Ty * c.diff_layers * (3.0 * DC * 2 * DC + (double)DC * 2 * DC) -
                           2.0 * B * (double)Ty * DC * DC;  // the last layer has no residual half
      ProfScope prof(PROF_DIFFNET, s, flops, 0.0);
      h->diffnet.run(r, s);
    }
    for (int l = 0; l < (fused ? 0 : c.diff_layers); ++l) {
      const DiffLayerW& lw = h->diff[l];
      const bool last = (l + 1 == c.diff_layers);
      // z = sigmoid(gate) * tanh(filter) of dilated_conv(h + step_emb) + cond_proj  (denoiser.py:69-77)
      auto d = conv_desc(lw.dilated, w.h, B, Ty, w.z);
      d.out_bs = bsD; d.out_ld = DC;
      d.addend = w.condp + (size_t)l * 2 * DC; d.addend_bs = (int64_t)Ty * CP; d.addend_ld = CP;
      d.act = PTTSPP_ACT_GATE;
      if (um) {
        planes_in(d, w.yh, w.yl, lw.dilated, 0);
        d.out = nullptr;
        planes_out(d, w.zh, w.zl, nullptr);
      } else {
        d.in_add = step_emb + (size_t)l * DC;
      }
      conv1d_cl(d, s);
      // residual half: h = (h + W_r z + b_r) / sqrt(2)   (denoiser.py:79-83)
      auto r = conv_desc(lw.outp, w.z, B, Ty, w.h);
      r.Cout = DC; r.out_bs = bsD; r.out_ld = DC;
      r.res = w.h; r.res_bs = bsD; r.res_ld = DC; r.out_div = sqrt2;
      if (um) {
        planes_in(r, w.zh, w.zl, lw.outp, 0);
        if (!last) planes_out(r, w.yh, w.yl, step_emb + (size_t)(l + 1) * DC);  // next layer's conv input
        if (h_planes) {
          // h = y - step_emb[l] from the planes (in place: every element is read and rewritten by the same thread)
          r.res = nullptr; r.out = nullptr;
          r.res_hi = w.yh; r.res_lo = w.yl; r.res_plane_sub = step_emb + (size_t)l * DC;
          r.res_plane_bs = bsD; r.res_plane_ld = DC;
        }
      } else {
        conv1d_cl(r, s);
      }
      // skip half: skip (+)= W_s z + b_s
      auto k = conv_desc(lw.outp, w.z, B, Ty, w.skip);
      k.w = lw.outp.w + DC; k.bias = lw.outp.bias + DC; k.Cout = DC; k.out_bs = bsD; k.out_ld = DC;
      k.beta = (l == 0) ? 0.f : 1.f;
      if (um) {
        planes_in(k, w.zh, w.zl, lw.outp, DC);
        if (last) planes_out(k, w.sh, w.sl, nullptr);  // operand planes of the skip sum
        if (last && h_planes) conv1d_cl(k, s);          // the last layer's residual half has no consumer
        else conv1d_umma_dual_cl(r, k, s);              // one launch, two epilogues (residual | skip)
      } else {
        conv1d_cl(k, s);
      }
    }
    const float sigma = (step > 0) ? expf(0.5f * h->logvar[step]) : 0.f;
    const float* z_step;
    if (z) {
      z_step = z + (size_t)(c.K_step - 1 - step) * B * M * Ty;
    } else {
      // the reference draws noise_like(x.shape) at every step, the last one included (diffusion.py:218)
      rng_offset += philox_normal(w.znoise, (int64_t)B * M * Ty, rng_seed, rng_offset, s);
      z_step = w.znoise;
    }
    if (tail) {
      DiffTailRun tr;
      tr.B = B; tr.T = Ty; tr.skip_hi = w.sh; tr.skip_lo = w.sl; tr.x = w.xt; tr.z = z_step;
      tr.c_recip = h->c_recip[step]; tr.c_recipm1 = h->c_recipm1[step]; tr.coef1 = h->coef1[step]; tr.coef2 = h->coef2[step];
      tr.sigma = sigma;
      tr.step_next = (step > 0) ? h->step_table + (size_t)(step - 1) * c.diff_layers * DC : nullptr;
      tr.y_hi = w.yh; tr.y_lo = w.yl;
      const double rows = (double)B * Ty;
      ProfScope prof(PROF_CONV_UMMA, s, rows * 2.0 * (DC * DC + (double)DC * M + (step > 0 ? (double)M * DC : 0.0)), 0.0);
      h->tail.run(tr, s);
      continue;
    }
    {
      auto d = conv_desc(h->skip_proj, w.skip, B, Ty, w.s);
      d.acc_scale = inv_sqrt_layers; d.act = PTTSPP_ACT_RELU;
      if (um) {
        planes_in(d, w.sh, w.sl, h->skip_proj, 0);
        d.out = nullptr;
        planes_out(d, w.ph, w.pl, nullptr);
      }
      conv1d_cl(d, s);
      auto e = conv_desc(h->out_proj, w.s, B, Ty, w.eps);
      if (um) planes_in(e, w.ph, w.pl, h->out_proj, 0);
      conv1d_cl(e, s);
    }
    ddpm_update(w.xt, w.eps, z_step, B, Ty, M, h->c_recip[step],
                h->c_recipm1[step], h->coef1[step], h->coef2[step], sigma, s, tc_inproj ? w.xh : nullptr,
                tc_inproj ? w.xl : nullptr, Mp);
  }
  // de-normalise, mask, back to [B][mel][Ty]  (diffusion.py:170-173, model.py:319-320)
  if (c.norm_scale > 0.f)
    transpose_btc_to_bct_affine(w.xt, mel, B, Ty, M, flen, c.norm_scale, 0.f, s);
  else
    transpose_btc_to_bct_affine(w.xt, mel, B, Ty, M, flen, 0.5f * (c.a_max - c.a_min), 0.5f * (c.a_max - c.a_min) + c.a_min, s);
  if (rng_offset_out) *rng_offset_out = rng_offset;
  PT_API_END
}

extern "C" int pttspp_acoustic_decode(pttspp_acoustic_t* h, const float* enc_state, const int64_t* dur,
                                      const int64_t* frame_len, int B, int Tx, int Ty, const float* pe_abs,
                                      const float* x_T, const float* z, float* mel, float* log_cf0, float* vuv,
                                      float* cond_out, void* workspace, size_t workspace_bytes,
                                      pttspp_stream_t stream) {
  if (!z) {
    pttspp::set_last_error("pttspp_acoustic_decode: z is NULL (use pttspp_acoustic_decode_rng to draw the noise natively)");
    return 1;
  }
  return acoustic_decode_impl(h, enc_state, dur, frame_len, B, Tx, Ty, pe_abs, x_T, z, 0, 0, nullptr, mel, log_cf0, vuv,
                              cond_out, workspace, workspace_bytes, stream);
}

extern "C" int pttspp_acoustic_decode_rng(pttspp_acoustic_t* h, const float* enc_state, const int64_t* dur,
                                          const int64_t* frame_len, int B, int Tx, int Ty, const float* pe_abs,
                                          const float* x_T, uint64_t rng_seed, uint64_t rng_offset,
                                          uint64_t* rng_offset_out, float* mel, float* log_cf0, float* vuv,
                                          float* cond_out, void* workspace, size_t workspace_bytes,
                                          pttspp_stream_t stream) {
  return acoustic_decode_impl(h, enc_state, dur, frame_len, B, Tx, Ty, pe_abs, x_T, nullptr, rng_seed, rng_offset,
                              rng_offset_out, mel, log_cf0, vuv, cond_out, workspace, workspace_bytes, stream);
}
