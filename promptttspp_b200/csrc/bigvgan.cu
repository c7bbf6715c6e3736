// BigVGAN generator forward (promptttspp/vocoders/bigvgan.py:71-131) on channels-last activations.
//
//   mel [B][80][T] -> transpose -> conv_pre (k7) -> 4 x { polyphase ConvTranspose1d,
//   3 AMP blocks x 3 layers of [AA-Snake -> dilated conv -> AA-Snake -> conv -> +x], averaged }
//   -> AA-Snake -> conv_post (k7, C->1) -> tanh -> wav [B][1][240 T]
//
// Weight-norm is folded at finalize(); the MRF average is accumulated in the epilogue of each
// block's last conv (out = (x + y)/n + out_old), so no separate add/div passes exist.
#include "common.h"

namespace pttspp {

struct AAParams {
  float* log_alpha = nullptr;
  float* up_f = nullptr;
  float* down_f = nullptr;
  int sym = 0;  // both filters exactly symmetric (f[k] == f[11-k]): the channel-pair AA kernel applies
};

struct AMPLayerW {
  PackedConv conv1, conv2;
  AAParams act1, act2;
};

struct UpsampleW {
  float* w = nullptr;  // [stride][J][Cin][w_ld]
  float* bias = nullptr;
  int Cin = 0, Cout = 0, Kt = 0, stride = 0, pad = 0, out_pad = 0, J = 0, w_ld = 0;
  // tcgen05 path: per phase r the 2-tap conv's split-fp16 weight planes [J][Cout][Cin] (+ its power-of-two scale)
  std::vector<void*> w_hi, w_lo;
  std::vector<float> w_scale_inv;
};

}  // namespace pttspp

struct pttspp_bigvgan {
  pttspp_bigvgan_config cfg;
  pttspp::TensorStore store;
  pttspp::DeviceBuffers dev;
  bool finalized = false;
  bool use_umma = true;  // PTTSPP_DISABLE_UMMA=1: every conv on the fp32 CUDA-core path
  // PTTSPP_AA_FUSE=1: the activation inside the conv kernel's producer warps on the 32 / 64-channel stages.  Bit-identical,
  // but measured SLOWER (48.0 vs 42.9 ms per cfg3 forward: the activation is FP32-pipe / latency bound and needs all 16
  // resident warps of an SM, the 7-8 producer warps of the fused kernel take 2.8x the stand-alone kernel's time --
  // profiles/r02_aa_conv_fusion_experiment.txt), so it stays opt-in.
  bool fuse_aa = false;
  int conv_impl = 4;  // 4: long contractions on the CTA-pair kernel, no chunked accumulation; PTTSPP_BIGVGAN_IMPL=2: round-2b choice
  pttspp::PackedConv conv_pre, conv_post;
  pttspp::PackedConv conv_pre_tc;  // tensor-core copy of conv_pre, mel axis zero-padded to a multiple of 64 (80 -> 128)
  std::vector<pttspp::UpsampleW> ups;
  std::vector<std::vector<std::vector<pttspp::AMPLayerW>>> mrfs;  // [stage][kernel][layer]
  pttspp::AAParams act_post;
  float* post_w = nullptr;  // conv_post weight as [K][C] for the dedicated C -> 1 kernel
  // F0-aware variant (vocoders/bigvgan_f0.py): per stage the Conv1d(1 -> C_i, k, stride) applied to the harmonic source
  struct NoiseConv {
    float* w = nullptr;  // [K][C]
    float* bias = nullptr;
    int K = 1, stride = 1, pad = 0, C = 0;
  };
  std::vector<NoiseConv> noise_convs;  // empty for the plain BigVGAN
};

namespace pttspp {
namespace {

AAParams load_aa(const TensorStore& st, DeviceBuffers& dev, const std::string& prefix, int C) {
  AAParams a;
  a.log_alpha = dev.upload(st.get(prefix + ".act.alpha", C).data);
  a.up_f = dev.upload(st.get(prefix + ".up.filter", 12).data);
  a.down_f = dev.upload(st.get(prefix + ".down.lowpass.filter", 12).data);
  const auto& uf = st.get(prefix + ".up.filter", 12).data;
  const auto& df = st.get(prefix + ".down.lowpass.filter", 12).data;
  a.sym = 1;
  for (int k = 0; k < 6; ++k)
    if (uf[k] != uf[11 - k] || df[k] != df[11 - k]) a.sym = 0;
  return a;
}

int stage_channels(const pttspp_bigvgan_config& c, int stage /* 0 = conv_pre output */) {
  return c.upsample_initial_channel >> stage;
}

int total_upsample(const pttspp_bigvgan_config& c) {
  int r = 1;
  for (int i = 0; i < c.num_upsamples; ++i) r *= c.upsample_rates[i];
  return r;
}

int64_t max_stage_elems(const pttspp_bigvgan_config& c, int B, int T) {
  int64_t mx = (int64_t)T * std::max(c.upsample_initial_channel, round_up(c.in_channel, 4));
  int64_t L = T;
  for (int i = 0; i < c.num_upsamples; ++i) {
    L *= c.upsample_rates[i];
    mx = std::max(mx, L * stage_channels(c, i + 1));
  }
  return mx * B;
}


// conv_post (C = 32 -> 1, k taps) + tanh, HBM bound (reads C floats per output sample): a warp owns a strip of
// output samples, lane = channel, the k-row window slides through registers, one shuffle reduction per sample and
// one coalesced 128-byte store per 32 samples.  Replaces bigvgan.py:129-131 for the 32-channel last stage; the generic
// implicit-GEMM kernel spent 1.7 ms on this N = 1 contraction (ncu launch list, profiles/).
constexpr int POST_STRIP = 128;  // outputs per warp
template <int K>
__global__ void __launch_bounds__(256) conv_post_tanh_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int L, int strips_per_row) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = warp / strips_per_row, t0 = (warp - b * strips_per_row) * POST_STRIP;
  if (t0 >= L) return;
  constexpr int P = K / 2;
  float wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wk[k] = w[k * 32 + lane];
  const float bz = bias ? bias[0] : 0.f;
  const float* xb = x + (int64_t)b * L * 32 + lane;
  float win[K];  // win[k] = x[t + k - P] for the current output t
#pragma unroll
  for (int k = 0; k < K - 1; ++k) {
    const int l = t0 + k - P;
    win[k + 1] = (l >= 0 && l < L) ? xb[(int64_t)l * 32] : 0.f;
  }
  float keep = 0.f;
  for (int tt = 0; tt < POST_STRIP; ++tt) {
    const int t = t0 + tt;
#pragma unroll
    for (int k = 0; k < K - 1; ++k) win[k] = win[k + 1];
    const int l = t + P;
    win[K - 1] = (l < L) ? xb[(int64_t)l * 32] : 0.f;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) acc = fmaf(win[k], wk[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tt & 31) == lane) keep = acc;
    if ((tt & 31) == 31) {
      const int to = t - 31 + lane;
      if (to < L) out[(int64_t)b * L + to] = tanhf(keep + bz);
    }
  }
}

void conv_post_tanh(const float* x, const float* w, const float* bias, float* out, int B, int L, int K, cudaStream_t s) {
  const int strips = ceil_div(L, POST_STRIP);
  const long long warps = (long long)B * strips;
  ProfScope prof(PROF_OTHER, s, 0.0, 4.0 * B * (double)L * 33);
  const unsigned blocks = (unsigned)ceil_div64(warps * 32, 256);
  PT_CHECK(K == 7, "conv_post kernel is instantiated for k = 7");
  conv_post_tanh_kernel<7><<<blocks, 256, 0, s>>>(x, w, bias, out, L, strips);
  PT_LAUNCHED();
}

// ---- F0-aware BigVGAN: harmonic-plus-noise source (vocoders/nsf.py) ------------------------------------------------
// SineGen._f02sine (:55-85) is two prefix sums over the sample axis per harmonic:
//   rad[t]   = ((h+1) * f0[t] / sr) mod 1            (+ rand_ini[h] at t = 0)
//   c1       = cumsum(rad); wrap[t] = (c1[t] mod 1) < (c1[t-1] mod 1)      (the "-1 whenever the sum passes 1" trick)
//   phase    = cumsum(rad - wrap);  sine = sin(2*pi*phase)
// torch's CPU cumsum accumulates in double and rounds every output to float; the kernels below do the same
// (chunked: per-chunk double sums, a tiny sequential scan of the chunk sums, then the in-chunk pass), so the result
// does not depend on the chunking beyond 1e-16.  One thread owns (utterance, chunk) and all harmonics.
constexpr int NSF_CHUNK = 256;
constexpr int NSF_MAXH = 16;

__device__ __forceinline__ float nsf_rad(float f0, int h, float sr, float ini, bool first) {
  const float fh = (h == 0) ? f0 : f0 * (float)(h + 1);  // f0_buf[:, :, idx+1] = f0_buf[:, :, 0] * (idx + 2)
  float r = fh / sr;
  r = r - floorf(r);  // python-style % 1 (non-negative operands)
  if (first) r = r + ini;
  return r;
}
__device__ __forceinline__ float nsf_mod1(float x) { return x - floorf(x); }

// pass 1 / 3: chunk sums.  shifted == 0: sum of rad; shifted == 1: sum of (rad + wrap shift), needs off1 (exclusive
// double prefix of pass 1).  sums / off: [B][nchunk][H]
__global__ void __launch_bounds__(128) nsf_chunk_sums_kernel(const float* __restrict__ f0, const float* __restrict__ rand_ini,
                                                             const double* __restrict__ off1, double* __restrict__ sums,
                                                             int B, int T, int hop, int H, float sr, int nchunk,
                                                             int shifted) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * nchunk) return;
  const int b = idx / nchunk, ch = idx - b * nchunk;
  const int L = T * hop, t0 = ch * NSF_CHUNK, t1 = min(L, t0 + NSF_CHUNK);
  for (int h = 0; h < H; ++h) {
    const float ini = rand_ini[b * H + h];
    double acc = 0.0, c1 = shifted ? off1[(size_t)idx * H + h] : 0.0;
    float prev = shifted ? nsf_mod1((float)c1) : 0.f;  // (c1[t0-1] mod 1); unused at t0 == 0
    for (int t = t0; t < t1; ++t) {
      const float r = nsf_rad(f0[b * T + t / hop], h, sr, ini, t == 0);
      if (!shifted) {
        acc += (double)r;
      } else {
        c1 += (double)r;
        const float cur = nsf_mod1((float)c1);
        const float sh = (t > 0 && (cur - prev) < 0.f) ? -1.f : 0.f;
        prev = cur;
        acc += (double)(r + sh);
      }
    }
    sums[(size_t)idx * H + h] = acc;
  }
}
// exclusive prefix of the chunk sums along the chunk axis (one thread per (b, h); nchunk is ~1000)
__global__ void nsf_chunk_scan_kernel(const double* __restrict__ sums, double* __restrict__ off, int B, int H, int nchunk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, h = idx - b * H;
  double acc = 0.0;
  for (int ch = 0; ch < nchunk; ++ch) {
    const size_t i = ((size_t)b * nchunk + ch) * H + h;
    off[i] = acc;
    acc += sums[i];
  }
}
// pass 5: sines, voiced/unvoiced mix with the injected noise (SineGen.forward :116-148), Linear(H -> 1) + tanh
// (SourceModuleHnNSF.forward :193-206) -> har_source [B][L]
__global__ void __launch_bounds__(128) nsf_source_kernel(const float* __restrict__ f0, const float* __restrict__ rand_ini,
                                                         const float* __restrict__ noise /*[B][L][H]*/,
                                                         const double* __restrict__ off1, const double* __restrict__ off2,
                                                         const float* __restrict__ lin_w, const float* __restrict__ lin_b,
                                                         float* __restrict__ har, int B, int T, int hop, int H, float sr,
                                                         float sine_amp, float noise_std, float voiced_thr, int nchunk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * nchunk) return;
  const int b = idx / nchunk, ch = idx - b * nchunk;
  const int L = T * hop, t0 = ch * NSF_CHUNK, t1 = min(L, t0 + NSF_CHUNK);
  double c1[NSF_MAXH], c2[NSF_MAXH];
  float prev[NSF_MAXH], ini[NSF_MAXH], lw[NSF_MAXH];
#pragma unroll
  for (int h = 0; h < NSF_MAXH; ++h) {
    if (h < H) {
      c1[h] = off1[(size_t)idx * H + h];
      c2[h] = off2[(size_t)idx * H + h];
      prev[h] = nsf_mod1((float)c1[h]);
      ini[h] = rand_ini[b * H + h];
      lw[h] = lin_w[h];
    }
  }
  const float two = 2.f, pi = 3.14159265358979323846f;  // `cumsum * 2 * np.pi`: two fp32 multiplications
  for (int t = t0; t < t1; ++t) {
    const float f = f0[b * T + t / hop];
    const float uv = (f > voiced_thr) ? 1.f : 0.f;
    const float namp = uv * noise_std + (1.f - uv) * sine_amp / 3.f;
    float acc = 0.f;
#pragma unroll
    for (int h = 0; h < NSF_MAXH; ++h) {
      if (h < H) {
        const float r = nsf_rad(f, h, sr, ini[h], t == 0);
        c1[h] += (double)r;
        const float cur = nsf_mod1((float)c1[h]);
        const float sh = (t > 0 && (cur - prev[h]) < 0.f) ? -1.f : 0.f;
        prev[h] = cur;
        c2[h] += (double)(r + sh);
        const float sine = sinf(((float)c2[h] * two) * pi) * sine_amp;
        const float sw = sine * uv + namp * noise[((size_t)b * L + t) * H + h];
        acc = fmaf(sw, lw[h], acc);  // F.linear: dot product over the harmonics
      }
    }
    har[(size_t)b * L + t] = tanhf(acc + lin_b[0]);
  }
}

// x_source = Conv1d(1 -> C, K, stride, pad)(har_source) written channels-last into the stage buffer, to which the
// transposed conv then ACCUMULATES (x = up(x) + noise_conv(har_source), bigvgan_f0.py:104-106)
__global__ void __launch_bounds__(256) noise_conv_kernel(const float* __restrict__ har, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int L,
                                                         int Lout, int C, int K, int stride, int pad) {
  extern __shared__ float win[];  // (rows-1)*stride + K samples
  constexpr int ROWS = 32;
  const int b = blockIdx.y, l0 = blockIdx.x * ROWS;
  const int span = (ROWS - 1) * stride + K;
  const int s0 = l0 * stride - pad;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int t = s0 + i;
    win[i] = (t >= 0 && t < L) ? har[(size_t)b * L + t] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ROWS * C; o += blockDim.x) {
    const int r = o / C, c = o - r * C;
    if (l0 + r >= Lout) continue;
    float acc = bias[c];
    for (int k = 0; k < K; ++k) acc = fmaf(win[r * stride + k], w[k * C + c], acc);
    out[((size_t)b * Lout + l0 + r) * C + c] = acc;
  }
}

}  // namespace
}  // namespace pttspp

using namespace pttspp;

extern "C" int pttspp_bigvgan_create(const pttspp_bigvgan_config* cfg, pttspp_bigvgan_t** out) {
  PT_API_BEGIN
  PT_CHECK(cfg && out, "null argument");
  PT_CHECK(cfg->num_upsamples >= 1 && cfg->num_upsamples <= 8 && cfg->num_kernels >= 1 && cfg->num_kernels <= 8 &&
               cfg->num_dilations >= 1 && cfg->num_dilations <= 8,
           "bigvgan: config out of range");
  PT_CHECK(cfg->in_channel % 16 == 0, "bigvgan: in_channel=%d must be a multiple of 16", cfg->in_channel);
  PT_CHECK((cfg->upsample_initial_channel >> cfg->num_upsamples) >= 16 &&
               (cfg->upsample_initial_channel >> cfg->num_upsamples) % 16 == 0,
           "bigvgan: channels must stay a multiple of 16 through all stages");
  for (int i = 0; i < cfg->num_upsamples; ++i)
    PT_CHECK(cfg->upsample_kernel_sizes[i] % cfg->upsample_rates[i] == 0,
             "bigvgan: upsample kernel %d must be a multiple of its rate %d", cfg->upsample_kernel_sizes[i],
             cfg->upsample_rates[i]);
  for (int j = 0; j < cfg->num_kernels; ++j)
    PT_CHECK(cfg->resblock_kernel_sizes[j] % 2 == 1, "bigvgan: resblock kernel sizes must be odd");
  auto* h = new pttspp_bigvgan();
  h->cfg = *cfg;
  *out = h;
  PT_API_END
}

extern "C" void pttspp_bigvgan_destroy(pttspp_bigvgan_t* h) { delete h; }

extern "C" int pttspp_bigvgan_set_tensor(pttspp_bigvgan_t* h, const char* name, const float* data,
                                         const int64_t* shape, int ndim, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(h, "null handle");
  h->store.set(name, data, shape, ndim, (cudaStream_t)stream);
  h->finalized = false;
  PT_API_END
}

extern "C" int pttspp_bigvgan_finalize(pttspp_bigvgan_t* h, pttspp_stream_t) {
  PT_API_BEGIN
  PT_CHECK(h, "null handle");
  const auto& c = h->cfg;
  h->dev.release();
  h->ups.clear();
  h->mrfs.clear();
  const int C0 = c.upsample_initial_channel;
  {
    const char* e = getenv("PTTSPP_DISABLE_UMMA");
    h->use_umma = !(e && e[0] == '1');
    const char* f = getenv("PTTSPP_AA_FUSE");
    h->fuse_aa = (f && f[0] == '1');
    const char* ci = getenv("PTTSPP_BIGVGAN_IMPL");
    h->conv_impl = (ci && ci[0] == '2') ? 2 : 4;
  }
  h->conv_pre = load_conv1d(h->store, h->dev, "conv_pre", C0, c.in_channel, 7, 1, 3);
  h->conv_pre_tc = PackedConv();
  if (h->use_umma && C0 % 64 == 0) {
    // conv_pre on the tensor cores (it was the last fp32 CUDA-core conv of the forward: 0.35 ms at cfg3): the mel axis is
    // zero-padded to a multiple of 64, which leaves the weight-norm over (Cin, K) unchanged
    const int Cp = round_up(c.in_channel, 64);
    const int64_t n = (int64_t)C0 * c.in_channel * 7;
    const bool wn = !h->store.has("conv_pre.weight");
    const auto& v = h->store.get(wn ? "conv_pre.weight_v" : "conv_pre.weight", n).data;  // [C0][in][7]
    std::vector<float> vp((size_t)C0 * Cp * 7, 0.f);
    for (int co = 0; co < C0; ++co)
      for (int ci = 0; ci < c.in_channel; ++ci)
        for (int k = 0; k < 7; ++k) vp[((size_t)co * Cp + ci) * 7 + k] = v[((size_t)co * c.in_channel + ci) * 7 + k];
    std::vector<uint16_t> hi(vp.size()), lo(vp.size());
    h->conv_pre_tc = h->conv_pre;
    h->conv_pre_tc.Cin = Cp;
    pack_conv_weight_split(vp.data(), wn ? h->store.get("conv_pre.weight_g", C0).data.data() : nullptr, C0, Cp, 7, hi.data(),
                           lo.data(), 0, &h->conv_pre_tc.w_scale_inv);
    h->conv_pre_tc.w_hi = h->dev.upload_bytes(hi.data(), hi.size() * 2);
    h->conv_pre_tc.w_lo = h->dev.upload_bytes(lo.data(), lo.size() * 2);
  }
  for (int i = 0; i < c.num_upsamples; ++i) {
    UpsampleW u;
    u.Cin = C0 >> i;
    u.Cout = C0 >> (i + 1);
    u.Kt = c.upsample_kernel_sizes[i];
    u.stride = c.upsample_rates[i];
    u.pad = u.stride / 2 + u.stride % 2;  // bigvgan.py:98
    u.out_pad = u.stride % 2;             // bigvgan.py:99
    u.J = u.Kt / u.stride;
    u.w_ld = round_up(u.Cout, 4);
    const std::string p = "upsamples." + std::to_string(i);
    std::vector<float> packed((size_t)u.stride * u.J * u.Cin * u.w_ld);
    const int64_t n = (int64_t)u.Cin * u.Cout * u.Kt;
    if (h->store.has(p + ".weight"))
      pack_convtr_weight(h->store.get(p + ".weight", n).data.data(), nullptr, u.Cin, u.Cout, u.Kt, u.stride,
                         packed.data(), u.w_ld, 0);
    else
      pack_convtr_weight(h->store.get(p + ".weight_v", n).data.data(), h->store.get(p + ".weight_g", u.Cin).data.data(),
                         u.Cin, u.Cout, u.Kt, u.stride, packed.data(), u.w_ld, 0);
    u.w = h->dev.upload(packed);
    u.bias = h->dev.upload(h->store.get(p + ".bias", u.Cout).data);
    if (h->use_umma && u.Cin % 64 == 0 && u.Cout % 16 == 0) {
      // per phase: the [J][Cin][Cout] taps as a torch-layout conv weight [Cout][Cin][J] -> split-fp16 planes
      std::vector<float> wt((size_t)u.Cout * u.Cin * u.J);
      std::vector<uint16_t> hi(wt.size()), lo(wt.size());
      for (int r = 0; r < u.stride; ++r) {
        for (int kp = 0; kp < u.J; ++kp)
          for (int ci = 0; ci < u.Cin; ++ci)
            for (int co = 0; co < u.Cout; ++co)
              wt[((size_t)co * u.Cin + ci) * u.J + kp] = packed[(((size_t)r * u.J + kp) * u.Cin + ci) * u.w_ld + co];
        float sc = 0.f;
        pack_conv_weight_split(wt.data(), nullptr, u.Cout, u.Cin, u.J, hi.data(), lo.data(), 0, &sc);
        u.w_hi.push_back(h->dev.upload_bytes(hi.data(), hi.size() * 2));
        u.w_lo.push_back(h->dev.upload_bytes(lo.data(), lo.size() * 2));
        u.w_scale_inv.push_back(sc);
      }
    }
    h->ups.push_back(u);

    std::vector<std::vector<AMPLayerW>> stage;
    for (int j = 0; j < c.num_kernels; ++j) {
      std::vector<AMPLayerW> block;
      const int k = c.resblock_kernel_sizes[j];
      for (int l = 0; l < c.num_dilations; ++l) {
        const int dl = c.resblock_dilations[j][l];
        const std::string lp = "mrfs." + std::to_string(i) + "." + std::to_string(j) + ".layers." + std::to_string(l);
        AMPLayerW w;
        w.conv1 = load_conv1d(h->store, h->dev, lp + ".conv1", u.Cout, u.Cout, k, dl, (k * dl - dl) / 2);
        w.conv2 = load_conv1d(h->store, h->dev, lp + ".conv2", u.Cout, u.Cout, k, 1, k / 2);
        if (h->use_umma && (u.Cout % 64 == 0 || u.Cout == 32)) {  // tensor-core path for the dense C -> C contractions
          attach_split_weights(h->store, h->dev, lp + ".conv1", w.conv1, false);
          attach_split_weights(h->store, h->dev, lp + ".conv2", w.conv2, false);
        }
        w.act1 = load_aa(h->store, h->dev, lp + ".act1", u.Cout);
        w.act2 = load_aa(h->store, h->dev, lp + ".act2", u.Cout);
        block.push_back(w);
      }
      stage.push_back(block);
    }
    h->mrfs.push_back(stage);
  }
  h->noise_convs.clear();
  if (h->store.has("noise_convs.0.weight")) {  // F0-aware generator (bigvgan_f0.py:66-79)
    for (int i = 0; i < c.num_upsamples; ++i) {
      pttspp_bigvgan::NoiseConv nc;
      nc.C = C0 >> (i + 1);
      int sf = 1;
      for (int q = i + 1; q < c.num_upsamples; ++q) sf *= c.upsample_rates[q];
      if (i + 1 < c.num_upsamples) { nc.K = 2 * sf; nc.stride = sf; nc.pad = sf / 2; }
      else { nc.K = 1; nc.stride = 1; nc.pad = 0; }
      const std::string p = "noise_convs." + std::to_string(i);
      const auto& wt = h->store.get(p + ".weight", (int64_t)nc.C * nc.K).data;  // [C][1][K]
      std::vector<float> kc((size_t)nc.K * nc.C);
      for (int co = 0; co < nc.C; ++co)
        for (int k = 0; k < nc.K; ++k) kc[(size_t)k * nc.C + co] = wt[(size_t)co * nc.K + k];
      nc.w = h->dev.upload(kc);
      nc.bias = h->dev.upload(h->store.get(p + ".bias", nc.C).data);
      h->noise_convs.push_back(nc);
    }
  }
  const int Cl = C0 >> c.num_upsamples;
  h->act_post = load_aa(h->store, h->dev, "act_post", Cl);
  h->conv_post = load_conv1d(h->store, h->dev, "conv_post", 1, Cl, 7, 1, 3);
  h->post_w = nullptr;
  if (Cl == 32) {  // [1][C][7] (weight or weight-norm pair, folded by load_conv1d into packed [K][C][w_ld]) -> [K][C]
    std::vector<float> host((size_t)7 * h->conv_post.Cin * h->conv_post.w_ld);
    PT_CUDA(cudaMemcpy(host.data(), h->conv_post.w, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<float> kc((size_t)7 * 32);
    for (int k = 0; k < 7; ++k)
      for (int ci = 0; ci < 32; ++ci) kc[(size_t)k * 32 + ci] = host[((size_t)k * 32 + ci) * h->conv_post.w_ld];
    h->post_w = h->dev.upload(kc);
  }
  h->finalized = true;
  PT_API_END
}

extern "C" size_t pttspp_bigvgan_workspace_bytes(const pttspp_bigvgan_t* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  const int64_t el = (max_stage_elems(h->cfg, B, T) + 63) / 64 * 64;
  return (size_t)el * 6 * sizeof(float) + 256;
}

static void bigvgan_forward_impl(pttspp_bigvgan_t* h, const float* mel, const float* har_source, int B, int T, float* wav,
                                 void* workspace, size_t workspace_bytes, pttspp_stream_t stream) {
  {
  PT_CHECK(h && mel && wav, "null argument");
  PT_CHECK((har_source != nullptr) == !h->noise_convs.empty(),
           "bigvgan: the F0-aware generator needs a harmonic source (forward_f0), the plain one must not get one");
  PT_CHECK(h->finalized, "bigvgan: finalize() has not been called after the last set_tensor()");
  PT_CHECK(B >= 1 && T >= 1, "bigvgan: empty input (B=%d, T=%d)", B, T);
  PT_CHECK(workspace && workspace_bytes >= pttspp_bigvgan_workspace_bytes(h, B, T), "bigvgan: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const auto& c = h->cfg;
  const int64_t el = (max_stage_elems(c, B, T) + 63) / 64 * 64;
  float* base = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float* R[6];
  for (int i = 0; i < 6; ++i) R[i] = base + (int64_t)i * el;
  float *bx = R[0], *bxs = R[1], *bA = R[2], *bB = R[3], *t1 = R[4], *t2 = R[5];

  // mel [B][C][T] -> [B][T][C]; conv_pre
  transpose_bct_to_btc(mel, t1, B, c.in_channel, T, s);
  // operand planes of the current stage input (for the tensor-core transposed convs) live in the t2 region, which
  // is free whenever a stage's last conv or conv_pre runs: hi = first half, lo = second half
  // (a FUSED last conv reads its fp32 input from t2, so its output planes go to the t1 region, unused in fused layers)
  float* plane_region = t2;
  auto stage_planes = [&](int64_t n_elems, uint16_t*& ph, uint16_t*& pl) {
    ph = reinterpret_cast<uint16_t*>(plane_region);
    pl = ph + n_elems;
  };
  {
    auto d = conv_desc(h->conv_pre, t1, B, T, bxs);
    if (h->conv_pre_tc.w_hi) {  // padded operand planes of the mel rows in the (still unused) bA region
      const int Cp = h->conv_pre_tc.Cin;
      uint16_t* mh = reinterpret_cast<uint16_t*>(bA);
      uint16_t* ml = mh + (size_t)B * T * Cp;
      split_f16_pad(t1, (int64_t)B * T, c.in_channel, Cp, mh, ml, s);
      d.Cin = Cp; d.in_ld = Cp; d.in_bs = (int64_t)T * Cp;
      d.in_hi = mh; d.in_lo = ml; d.w_hi = h->conv_pre_tc.w_hi; d.w_lo = h->conv_pre_tc.w_lo;
      d.w_scale_inv = h->conv_pre_tc.w_scale_inv; d.impl = h->conv_impl;
    }
    if (!h->ups[0].w_hi.empty()) {
      uint16_t *ph, *pl;
      stage_planes((int64_t)B * T * h->conv_pre.Cout, ph, pl);
      d.out_hi = ph; d.out_lo = pl; d.out_plane_bs = (int64_t)T * h->conv_pre.Cout; d.out_plane_ld = h->conv_pre.Cout;
    }
    conv1d_cl(d, s);
  }
  int L = T;
  float* hcur = bxs;  // stage input [B][L][C]
  for (int i = 0; i < c.num_upsamples; ++i) {
    const UpsampleW& u = h->ups[i];
    const int Lout = (L - 1) * u.stride - 2 * u.pad + u.Kt + u.out_pad;
    const bool f0_aware = har_source != nullptr;
    if (f0_aware) {
      const auto& nc = h->noise_convs[i];
      const int Lh = T * total_upsample(c);
      dim3 grid(ceil_div(Lout, 32), B);
      const size_t sm = ((size_t)31 * nc.stride + nc.K) * sizeof(float);
      noise_conv_kernel<<<grid, 256, sm, s>>>(har_source, nc.w, nc.bias, bx, Lh, Lout, nc.C, nc.K, nc.stride, nc.pad);
      PT_LAUNCHED();
    }
    // polyphase transposed conv: phase r writes rows m*stride + r - pad
    for (int r = 0; r < u.stride; ++r) {
      pttspp_conv1d_desc d;
      memset(&d, 0, sizeof(d));
      d.in = hcur; d.in_bs = (int64_t)L * u.Cin; d.in_ld = u.Cin; d.T_in = L; d.Cin = u.Cin;
      d.w = u.w + (size_t)r * u.J * u.Cin * u.w_ld; d.w_ld = u.w_ld; d.bias = u.bias;
      d.K = u.J; d.dil = 1; d.pad = u.J - 1; d.in_stride = 1;
      d.out = bx; d.out_bs = (int64_t)Lout * u.Cout; d.out_ld = u.Cout; d.T_out = Lout; d.Cout = u.Cout;
      const int off = r - u.pad;
      d.m_begin = off < 0 ? ceil_div(-off, u.stride) : 0;
      const int m_end = (Lout - 1 - off) / u.stride;  // last m with m*stride + off <= Lout-1
      d.M = m_end - d.m_begin + 1;
      d.out_mul = u.stride; d.out_off = off;
      d.acc_scale = 1.f; d.res_scale = 1.f; d.alpha = 1.f; d.beta = f0_aware ? 1.f : 0.f; d.B = B;
      if (!u.w_hi.empty()) {
        uint16_t *ph, *pl;
        stage_planes((int64_t)B * L * u.Cin, ph, pl);
        d.in_hi = ph; d.in_lo = pl; d.w_hi = u.w_hi[r]; d.w_lo = u.w_lo[r]; d.w_scale_inv = u.w_scale_inv[r]; d.impl = h->conv_impl;
      }
      conv1d_cl(d, s);
    }
    L = Lout;
    const int C = u.Cout;
    for (int j = 0; j < c.num_kernels; ++j) {
      const float* cur = bx;
      for (int l = 0; l < c.num_dilations; ++l) {
        const AMPLayerW& w = h->mrfs[i][j][l];
        const bool um = w.conv1.w_hi != nullptr;
        // tcgen05 path: the activation writes its result as split-fp16 operand planes into the t1 region
        // (hi = first half, lo = second half: 2 x 2 bytes per element, the same footprint as fp32)
        uint16_t* ph = reinterpret_cast<uint16_t*>(t1);
        uint16_t* pl = ph + (size_t)B * L * C;
        auto use_planes = [&](pttspp_conv1d_desc& q, const PackedConv& pc) {
          q.in_hi = ph; q.in_lo = pl; q.w_hi = pc.w_hi; q.w_lo = pc.w_lo; q.w_scale_inv = pc.w_scale_inv; q.impl = h->conv_impl;
        };
        // narrow stages (32 / 64 channels): the activation runs inside the conv kernel's producer warps, the activated
        // tensor never exists in HBM (conv1d_umma.cu: aa_conv_wres_kernel); bit-identical to the two-launch path
        auto fused_desc = [&](const PackedConv& pc, const float* in, float* out) {
          auto q = conv_desc(pc, in, B, L, out);
          q.w_hi = pc.w_hi; q.w_lo = pc.w_lo; q.w_scale_inv = pc.w_scale_inv; q.impl = h->conv_impl;
          return q;
        };
        bool fuse1 = false, fuse2 = false;
        if (um && h->fuse_aa && w.act1.sym && w.act2.sym) {
          fuse1 = aa_conv1d_supported(fused_desc(w.conv1, cur, t2));
          fuse2 = aa_conv1d_supported(fused_desc(w.conv2, t2, t1));
        }
        if (fuse1) {
          aa_conv1d_cl(fused_desc(w.conv1, cur, t2), w.act1.log_alpha, w.act1.up_f, w.act1.down_f, s);
        } else {
          if (um) aa_snake_cl(cur, nullptr, B, L, C, w.act1.log_alpha, w.act1.up_f, w.act1.down_f, s, ph, pl, w.act1.sym);
          else aa_snake_cl(cur, t1, B, L, C, w.act1.log_alpha, w.act1.up_f, w.act1.down_f, s, nullptr, nullptr, w.act1.sym);
          auto d = conv_desc(w.conv1, t1, B, L, t2);
          if (um) use_planes(d, w.conv1);
          conv1d_cl(d, s);
        }
        if (!fuse2) {
          if (um) aa_snake_cl(t2, nullptr, B, L, C, w.act2.log_alpha, w.act2.up_f, w.act2.down_f, s, ph, pl, w.act2.sym);
          else aa_snake_cl(t2, t1, B, L, C, w.act2.log_alpha, w.act2.up_f, w.act2.down_f, s, nullptr, nullptr, w.act2.sym);
        }
        const bool last = (l == c.num_dilations - 1);
        float* dst = last ? bxs : ((cur == bA) ? bB : bA);
        auto d = fuse2 ? fused_desc(w.conv2, t2, dst) : conv_desc(w.conv2, t1, B, L, dst);
        if (um && !fuse2) use_planes(d, w.conv2);
        d.res = cur; d.res_bs = (int64_t)L * C; d.res_ld = C;
        if (last) {  // xs = (j ? xs : 0) + (x + y); the last block divides by num_kernels (bigvgan.py:124-127)
          d.beta = (j == 0) ? 0.f : 1.f;
          if (j == c.num_kernels - 1) {
            d.out_div = (float)c.num_kernels;
            if (i + 1 < c.num_upsamples && !h->ups[i + 1].w_hi.empty()) {
              // the finished stage output also leaves as operand planes for the next transposed conv
              uint16_t *oh, *ol;
              plane_region = fuse2 ? t1 : t2;
              stage_planes((int64_t)B * L * C, oh, ol);
              d.out_hi = oh; d.out_lo = ol; d.out_plane_bs = (int64_t)L * C; d.out_plane_ld = C;
            }
          }
        }
        if (fuse2) aa_conv1d_cl(d, w.act2.log_alpha, w.act2.up_f, w.act2.down_f, s);
        else conv1d_cl(d, s);
        cur = dst;
      }
    }
    hcur = bxs;
    // the next stage's transposed conv reads bxs and writes bx: no aliasing
  }
  const int Cl = c.upsample_initial_channel >> c.num_upsamples;
  aa_snake_cl(hcur, t1, B, L, Cl, h->act_post.log_alpha, h->act_post.up_f, h->act_post.down_f, s, nullptr, nullptr,
              h->act_post.sym);
  if (h->post_w && Cl == 32) {
    conv_post_tanh(t1, h->post_w, h->conv_post.bias, wav, B, L, 7, s);
  } else {
    auto d = conv_desc(h->conv_post, t1, B, L, wav);
    d.act = PTTSPP_ACT_TANH;
    conv1d_cl(d, s);
  }
  }
}

extern "C" int pttspp_bigvgan_forward(pttspp_bigvgan_t* h, const float* mel, int B, int T, float* wav, void* workspace,
                                      size_t workspace_bytes, pttspp_stream_t stream) {
  PT_API_BEGIN
  bigvgan_forward_impl(h, mel, nullptr, B, T, wav, workspace, workspace_bytes, stream);
  PT_API_END
}

extern "C" int pttspp_bigvgan_forward_f0(pttspp_bigvgan_t* h, const float* mel, const float* har_source, int B, int T,
                                         float* wav, void* workspace, size_t workspace_bytes, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(har_source, "null harmonic source");
  bigvgan_forward_impl(h, mel, har_source, B, T, wav, workspace, workspace_bytes, stream);
  PT_API_END
}

extern "C" size_t pttspp_nsf_source_workspace_bytes(int B, int T, int hop, int harmonic_num) {
  if (B <= 0 || T <= 0 || hop <= 0 || harmonic_num < 0) return 0;
  const size_t nchunk = ((size_t)T * hop + NSF_CHUNK - 1) / NSF_CHUNK;
  return 4 * (size_t)B * nchunk * (harmonic_num + 1) * sizeof(double) + 256;
}

extern "C" int pttspp_nsf_source(const float* f0, int B, int T, int hop, float sampling_rate, int harmonic_num,
                                 float sine_amp, float noise_std, float voiced_threshold, const float* rand_ini,
                                 const float* noise, const float* lin_w, const float* lin_b, float* har_source,
                                 void* workspace, size_t workspace_bytes, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(f0 && rand_ini && noise && lin_w && lin_b && har_source && workspace, "null argument");
  const int H = harmonic_num + 1;
  PT_CHECK(B >= 1 && T >= 1 && hop >= 1 && H >= 1 && H <= NSF_MAXH, "nsf_source: bad shape (harmonics <= %d)", NSF_MAXH - 1);
  PT_CHECK(workspace_bytes >= pttspp_nsf_source_workspace_bytes(B, T, hop, harmonic_num), "nsf_source: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int nchunk = ceil_div(T * hop, NSF_CHUNK);
  const size_t n = (size_t)B * nchunk * H;
  double* sums1 = (double*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  double *off1 = sums1 + n, *sums2 = off1 + n, *off2 = sums2 + n;
  const int threads = B * nchunk;
  ProfScope prof(PROF_OTHER, s, 0.0, 4.0 * B * (double)T * hop * (H + 1));
  nsf_chunk_sums_kernel<<<ceil_div(threads, 128), 128, 0, s>>>(f0, rand_ini, nullptr, sums1, B, T, hop, H, sampling_rate,
                                                                nchunk, 0);
  PT_LAUNCHED();
  nsf_chunk_scan_kernel<<<ceil_div(B * H, 64), 64, 0, s>>>(sums1, off1, B, H, nchunk);
  PT_LAUNCHED();
  nsf_chunk_sums_kernel<<<ceil_div(threads, 128), 128, 0, s>>>(f0, rand_ini, off1, sums2, B, T, hop, H, sampling_rate,
                                                                nchunk, 1);
  PT_LAUNCHED();
  nsf_chunk_scan_kernel<<<ceil_div(B * H, 64), 64, 0, s>>>(sums2, off2, B, H, nchunk);
  PT_LAUNCHED();
  nsf_source_kernel<<<ceil_div(threads, 128), 128, 0, s>>>(f0, rand_ini, noise, off1, off2, lin_w, lin_b, har_source, B, T,
                                                            hop, H, sampling_rate, sine_amp, noise_std, voiced_threshold,
                                                            nchunk);
  PT_LAUNCHED();
  PT_API_END
}
