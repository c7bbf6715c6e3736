// Small HBM-bound kernels of the acoustic model: layout changes, gathers (embedding, length
// regulator), MDN heads, the Conformer conv-module middle (GLU -> depthwise -> BatchNorm -> Swish),
// pitch embedding and the DDPM posterior update.  All use one thread per output element with the
// channel index fastest (coalesced 128-byte rows) or one warp per row with shuffle reductions.
#include <cuda_fp16.h>

#include "common.h"

namespace pttspp {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- layout ------------------------------------------------------------------------------
// in [B][R][S] -> out [B][S][R]  (generic 32x32 tile transpose); optional length mask on the
// dimension that ends up contiguous == `S` index meaning depends on caller, see wrappers.
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R,
                                                        int S, const int64_t* __restrict__ len, int len_on_r,
                                                        float scale, float shift) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
  const float* ib = in + (int64_t)b * R * S;
  float* ob = out + (int64_t)b * R * S;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, s = s0 + tx;
    tile[i][tx] = (r < R && s < S) ? ib[(int64_t)r * S + s] : 0.f;
  }
  __syncthreads();
  long long l = 0x7fffffffffffffffLL;
  if (len) l = len[b];
  for (int i = ty; i < 32; i += 8) {
    const int s = s0 + i, r = r0 + tx;
    if (r < R && s < S) {
      const int tpos = len_on_r ? r : s;
      ob[(int64_t)s * R + r] = ((long long)tpos < l) ? tile[tx][i] * scale + shift : 0.f;
    }
  }
}

// ---- embedding ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embedding_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ len,
                                                        const float* __restrict__ table, int T, int C, int vocab,
                                                        float scale1, float scale2, float* __restrict__ out,
                                                        int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t bt = i / C;
  const int t = (int)(bt % T), b = (int)(bt / T);
  float v = 0.f;
  if (!len || (long long)t < (long long)len[b]) {
    long long id = ids[bt];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    v = table[id * C + c] * scale1 * scale2;
  }
  out[i] = v;
}

// ---- Conformer conv module middle ---------------------------------------------------------
// in: [B][T][2C] = pointwise_conv1 output (already * mask).  GLU over the channel halves,
// depthwise conv (zero padding), * mask, eval-mode BatchNorm as y*scale + shift, Swish.
__global__ void __launch_bounds__(256) glu_dw_bn_swish_kernel(const float* __restrict__ in,
                                                              const int64_t* __restrict__ len,
                                                              const float* __restrict__ dw_w,
                                                              const float* __restrict__ dw_b,
                                                              const float* __restrict__ bn_scale,
                                                              const float* __restrict__ bn_shift, int T, int C, int K,
                                                              float* __restrict__ out, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t bt = i / C;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const float* ib = in + (int64_t)b * T * 2 * C;
  float acc = dw_b[c];
  const int half = (K - 1) / 2;
  for (int k = 0; k < K; ++k) {
    const int tt = t + k - half;
    if (tt >= 0 && tt < T) {
      const float a = ib[(int64_t)tt * 2 * C + c];
      const float g = ib[(int64_t)tt * 2 * C + C + c];
      acc = fmaf(dw_w[c * K + k], a * (1.f / (1.f + expf(-g))), acc);
    }
  }
  const float mask = (!len || (long long)t < (long long)len[b]) ? 1.f : 0.f;
  float y = (acc * mask) * bn_scale[c] + bn_shift[c];
  out[i] = y / (1.f + expf(-y));
}

// ---- style path ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_rows_kernel(float* __restrict__ x, int rows, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* xr = x + (int64_t)row * C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) ss = fmaf(xr[c], xr[c], ss);
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < C; c += 32) xr[c] = xr[c] / denom;
}

// one block per batch row; D threads-strided.  log_softmax over G per dimension, first arg-max wins.
__global__ void __launch_bounds__(256) style_mdn_sample_kernel(const float* __restrict__ logpi,
                                                               const float* __restrict__ logsigma,
                                                               const float* __restrict__ mu,
                                                               const float* __restrict__ z, int G, int D,
                                                               float noise_scale, int normalize,
                                                               float* __restrict__ style,
                                                               const float* __restrict__ comp_u) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* lp = logpi + (int64_t)b * G * D;
  const float* ls = logsigma + (int64_t)b * G * D;
  const float* m = mu + (int64_t)b * G * D;
  float ss = 0.f;
  for (int dd = threadIdx.x; dd < D; dd += blockDim.x) {
    float mx = -INFINITY;
    for (int g = 0; g < G; ++g) mx = fmaxf(mx, lp[g * D + dd]);
    float se = 0.f;
    for (int g = 0; g < G; ++g) se += expf(lp[g * D + dd] - mx);
    const float lse = logf(se);
    int best = 0;
    float bestv = -INFINITY;
    for (int g = 0; g < G; ++g) {
      const float v = (lp[g * D + dd] - mx) - lse;
      if (v > bestv) {
        bestv = v;
        best = g;
      }
    }
    if (comp_u) {
      // use_max=False (mdn.py:226-257): one Categorical(probs=exp(log_pi)) draw per dimension, here by inverse CDF of
      // the supplied uniform (the draw itself is an input, like every other random number of the path)
      const float u = comp_u[(int64_t)b * D + dd];
      float tot = 0.f;
      for (int g = 0; g < G; ++g) tot += expf((lp[g * D + dd] - mx) - lse);
      float cum = 0.f;
      best = G - 1;
      for (int g = 0; g < G; ++g) {
        cum += expf((lp[g * D + dd] - mx) - lse) / tot;
        if (u < cum) {
          best = g;
          break;
        }
      }
    }
    const float sigma = expf(ls[best * D + dd]);
    const float val = m[best * D + dd] + sigma * z[(int64_t)b * D + dd] * noise_scale;
    style[(int64_t)b * D + dd] = val;
    ss = fmaf(val, val, ss);
  }
  if (!normalize) return;
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float denom = fmaxf(sqrtf(red[0]), 1e-12f);
  for (int dd = threadIdx.x; dd < D; dd += blockDim.x) style[(int64_t)b * D + dd] /= denom;
}

__global__ void __launch_bounds__(256) add_row_broadcast_kernel(float* __restrict__ x, const float* __restrict__ v,
                                                                int T, int C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int b = (int)(i / ((int64_t)T * C));
  x[i] += v[(int64_t)b * C + c];
}

// ---- MDN duration head: warp per phoneme row ----------------------------------------------
constexpr int MAX_G = 16;
__global__ void __launch_bounds__(256) mdn_duration_head_kernel(const float* __restrict__ h,
                                                                const float* __restrict__ w_pi,
                                                                const float* __restrict__ b_pi,
                                                                const float* __restrict__ w_ls,
                                                                const float* __restrict__ b_ls,
                                                                const float* __restrict__ w_mu,
                                                                const float* __restrict__ b_mu, int rows, int C,
                                                                int G, float* __restrict__ log_d) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* hr = h + (int64_t)row * C;
  float pi[MAX_G], ls[MAX_G], mu[MAX_G];
#pragma unroll
  for (int g = 0; g < MAX_G; ++g) pi[g] = ls[g] = mu[g] = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float x = hr[c];
#pragma unroll
    for (int g = 0; g < MAX_G; ++g)
      if (g < G) {
        pi[g] = fmaf(x, w_pi[g * C + c], pi[g]);
        ls[g] = fmaf(x, w_ls[g * C + c], ls[g]);
        mu[g] = fmaf(x, w_mu[g * C + c], mu[g]);
      }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int g = 0; g < MAX_G; ++g)
    if (g < G) {
      pi[g] = warp_sum(pi[g]) + b_pi[g];
      ls[g] = warp_sum(ls[g]) + b_ls[g];
      mu[g] = warp_sum(mu[g]) + b_mu[g];
      mx = fmaxf(mx, pi[g]);
    }
  float se = 0.f;
#pragma unroll
  for (int g = 0; g < MAX_G; ++g)
    if (g < G) se += expf(pi[g] - mx);
  const float lse = logf(se);
  float bestv = -INFINITY, sig = 0.f, m = 0.f;
#pragma unroll
  for (int g = 0; g < MAX_G; ++g)
    if (g < G) {
      const float v = (pi[g] - mx) - lse;
      if (v > bestv) {
        bestv = v;
        sig = ls[g];
        m = mu[g];
      }
    }
  if (lane == 0) {
    const float sigma = expf(sig);
    log_d[row] = m + fmaxf(sigma * sigma, 1e-14f) / 2.f;
  }
}

// ---- duration quantisation + prefix sum (one CTA per utterance) ---------------------------
__global__ void __launch_bounds__(256) duration_quantize_kernel(const float* __restrict__ log_d,
                                                                const int64_t* __restrict__ phone_len, int Tx,
                                                                int64_t* __restrict__ dur,
                                                                int64_t* __restrict__ frame_len) {
  __shared__ long long part[256];
  const int b = blockIdx.x;
  const long long len = phone_len ? phone_len[b] : (long long)Tx;
  long long local = 0;
  for (int i = threadIdx.x; i < Tx; i += blockDim.x) {
    long long dq = 0;
    if ((long long)i < len) {
      // exp in double, rounded once to fp32: the correctly rounded value (a 1-ulp deviation of expf can flip
      // the integer at a .5 boundary and shift every later frame); rintf = round-half-to-even = torch.round
      const float e = rintf((float)exp((double)log_d[(int64_t)b * Tx + i]));
      dq = (long long)fmaxf(e, 1.f);
    }
    dur[(int64_t)b * Tx + i] = dq;
    local += dq;
  }
  part[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) frame_len[b] = part[0];
}

// ---- length regulator: out[b][t][:] = x[b][searchsorted(cumsum(dur[b]), t, right)][:] -------
constexpr int LR_ROWS = 32;  // frames per CTA
__global__ void __launch_bounds__(256) length_regulate_kernel(const float* __restrict__ x,
                                                              const int64_t* __restrict__ dur, int Tx, int C, int Ty,
                                                              float* __restrict__ out, int32_t* __restrict__ idx_out) {
  extern __shared__ int cum[];  // [Tx] inclusive prefix sums (frames fit int32)
  __shared__ int part[256];
  __shared__ int rows_idx[LR_ROWS];
  const int b = blockIdx.y, t0 = blockIdx.x * LR_ROWS;
  const int64_t* db = dur + (int64_t)b * Tx;
  // chunked scan: thread i owns elements [i*per, (i+1)*per)
  const int per = (Tx + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(lo + per, Tx);
  int s = 0;
  for (int i = lo; i < hi; ++i) {
    s += (int)db[i];
    cum[i] = s;
  }
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) {
      const int v = part[i];
      part[i] = run;
      run += v;
    }
  }
  __syncthreads();
  const int off = part[threadIdx.x];
  for (int i = lo; i < hi; ++i) cum[i] += off;
  __syncthreads();
  const int total = Tx > 0 ? cum[Tx - 1] : 0;
  if (threadIdx.x < LR_ROWS) {
    const int t = t0 + threadIdx.x;
    int idx = -1;
    if (t < Ty && t < total) {
      int l = 0, r = Tx;  // first i with cum[i] > t
      while (l < r) {
        const int mid = (l + r) >> 1;
        if (cum[mid] > t) r = mid; else l = mid + 1;
      }
      idx = l;
    }
    rows_idx[threadIdx.x] = idx;
    if (idx_out && t < Ty) idx_out[(int64_t)b * Ty + t] = idx;
  }
  __syncthreads();
  const int c4n = C >> 2;
  for (int e = threadIdx.x; e < LR_ROWS * c4n; e += blockDim.x) {
    const int r = e / c4n, c4 = e % c4n;
    const int t = t0 + r;
    if (t >= Ty) continue;
    const int idx = rows_idx[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx >= 0) v = *reinterpret_cast<const float4*>(x + ((int64_t)b * Tx + idx) * C + c4 * 4);
    *reinterpret_cast<float4*>(out + ((int64_t)b * Ty + t) * C + c4 * 4) = v;
  }
}

// ---- pitch head (Conv1d C->2, k=1) and pitch embedding ------------------------------------
__global__ void __launch_bounds__(256) pitch_head_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         const int64_t* __restrict__ len, int B, int T, int C,
                                                         float* __restrict__ log_cf0, float* __restrict__ vuv) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (int64_t)B * T) return;
  const int b = (int)(row / T), t = (int)(row % T);
  const float* hr = h + row * C;
  float a0 = 0.f, a1 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float x = hr[c];
    a0 = fmaf(x, w[c], a0);
    a1 = fmaf(x, w[C + c], a1);
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (lane == 0) {
    const float mask = (!len || (long long)t < (long long)len[b]) ? 1.f : 0.f;
    log_cf0[row] = (a0 + bias[0]) * mask;
    vuv[row] = (a1 + bias[1]) * mask;
  }
}

__global__ void __launch_bounds__(256) pitch_embed_add_kernel(float* __restrict__ x, const float* __restrict__ log_cf0,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              const int64_t* __restrict__ len, int T, int C,
                                                              int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t bt = i / C;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const float mask = (!len || (long long)t < (long long)len[b]) ? 1.f : 0.f;
  x[i] += (fmaf(w[c], log_cf0[bt], bias[c])) * mask;
}

// ---- DDPM ancestral step (diffusion.py:181-221) -------------------------------------------
// x, eps: [B][T][M] channels-last; z: [B][M][T] (the layout torch.randn drew it in).
// xp_hi/xp_lo (optional): split-fp16 planes [B][T][Mp] of the updated x for the tensor-core input projection of the
// next step (columns >= M stay zero: the planes are cleared once per call)
__global__ void __launch_bounds__(256) ddpm_update_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                                          const float* __restrict__ z, int T, int M, float c_recip,
                                                          float c_recipm1, float coef1, float coef2, float sigma,
                                                          __half* __restrict__ xp_hi, __half* __restrict__ xp_lo, int Mp) {
  __shared__ float zt[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* zb = z + (int64_t)b * M * T;
  for (int i = ty; i < 32; i += 8) {  // i: channel, tx: time (contiguous in z)
    const int m = m0 + i, t = t0 + tx;
    zt[i][tx] = (m < M && t < T) ? zb[(int64_t)m * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {  // i: time, tx: channel (contiguous in x)
    const int t = t0 + i, m = m0 + tx;
    if (t < T && m < M) {
      const int64_t idx = ((int64_t)b * T + t) * M + m;
      const float xv = x[idx];
      float x0 = c_recip * xv - c_recipm1 * eps[idx];
      x0 = fminf(fmaxf(x0, -1.f), 1.f);
      const float mean = coef1 * x0 + coef2 * xv;
      const float xn = mean + sigma * zt[tx][i];
      x[idx] = xn;
      if (xp_hi) {
        const int64_t pidx = ((int64_t)b * T + t) * Mp + m;
        const __half hh = pt_f2h_sat(xn);
        xp_hi[pidx] = hh;
        xp_lo[pidx] = pt_f2h_sat(xn - __half2float(hh));
      }
    }
  }
}


// ---- zero-phase IIR filter of the f0 contour (utils/model.py:164-196 -> torchaudio.functional.filtfilt) -------------
// filtfilt(x, a, b, clamp=False) = flip(lfilter(flip(lfilter(x)))) with zero initial state and no edge padding
// (torchaudio's, not scipy's, convention).  The recursion is sequential in time; rows (utterances) are independent and
// short (a few thousand 10-ms frames), so one thread owns one row and a ring of the last NT-1 inputs / outputs.
constexpr int IIR_MAX_TAPS = 8;
__global__ void __launch_bounds__(64) iir_filtfilt_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          float* __restrict__ tmp, int rows, int Tfull,
                                                          const float* __restrict__ bc, const float* __restrict__ ac, int nt,
                                                          const int64_t* __restrict__ row_len, int min_len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  // ragged batch: row r holds row_len[r] samples; both recursions run over exactly those (the backward pass starts at
  // the row's own last sample with zero state, as a per-utterance call would) and the padding is copied through
  int T = Tfull;
  if (row_len) {
    T = (int)min((int64_t)Tfull, max((int64_t)0, row_len[r]));
    const float* xs = x + (int64_t)r * Tfull;
    float* ys = y + (int64_t)r * Tfull;
    const int keep_from = (T <= min_len) ? 0 : T;  // rows too short for the filter pass through unchanged
    for (int n = keep_from; n < Tfull; ++n) ys[n] = xs[n];
    if (T <= min_len) return;
  }
  float b[IIR_MAX_TAPS], a[IIR_MAX_TAPS];
  const float a0 = ac[0];
#pragma unroll
  for (int k = 0; k < IIR_MAX_TAPS; ++k) {
    b[k] = (k < nt) ? bc[k] / a0 : 0.f;  // lfilter normalises both coefficient sets by a[0]
    a[k] = (k < nt) ? ac[k] / a0 : 0.f;
  }
  const float* xr = x + (int64_t)r * Tfull;
  float* t1 = tmp + (int64_t)r * Tfull;
  float* yr = y + (int64_t)r * Tfull;
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: forward over x -> t1; pass 1: backward over t1 -> y (a forward filter of the flipped signal)
    float xin[IIR_MAX_TAPS], yo[IIR_MAX_TAPS];
#pragma unroll
    for (int k = 0; k < IIR_MAX_TAPS; ++k) xin[k] = yo[k] = 0.f;
    for (int n = 0; n < T; ++n) {
      const int idx = pass ? (T - 1 - n) : n;
      const float v = pass ? t1[idx] : xr[idx];
#pragma unroll
      for (int k = IIR_MAX_TAPS - 1; k > 0; --k) {
        xin[k] = xin[k - 1];
        yo[k] = yo[k - 1];
      }
      xin[0] = v;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < IIR_MAX_TAPS; ++k) acc = fmaf(b[k], xin[k], acc);
#pragma unroll
      for (int k = 1; k < IIR_MAX_TAPS; ++k) acc = fmaf(-a[k], yo[k], acc);
      yo[0] = acc;
      if (pass) yr[idx] = acc;
      else t1[idx] = acc;
    }
  }
}

}  // namespace

static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)ceil_div64(n, per); }

void transpose_bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t s) {
  if (B == 0 || C == 0 || T == 0) return;
  dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B);
  transpose_kernel<<<grid, 256, 0, s>>>(in, out, C, T, nullptr, 0, 1.f, 0.f);
  PT_LAUNCHED();
}

void transpose_btc_to_bct(const float* in, float* out, int B, int T, int C, const int64_t* len, float scale,
                          cudaStream_t s) {
  if (B == 0 || C == 0 || T == 0) return;
  dim3 grid(ceil_div(C, 32), ceil_div(T, 32), B);
  // R = T (rows of the input), S = C; the length mask applies to the R index (time)
  transpose_kernel<<<grid, 256, 0, s>>>(in, out, T, C, len, 1, scale, 0.f);
  PT_LAUNCHED();
}

void transpose_btc_to_bct_affine(const float* in, float* out, int B, int T, int C, const int64_t* len, float scale,
                                 float shift, cudaStream_t s) {
  if (B == 0 || C == 0 || T == 0) return;
  dim3 grid(ceil_div(C, 32), ceil_div(T, 32), B);
  transpose_kernel<<<grid, 256, 0, s>>>(in, out, T, C, len, 1, scale, shift);
  PT_LAUNCHED();
}

void embedding_cl(const int64_t* ids, const int64_t* len, const float* table, int B, int T, int C, int vocab,
                  float scale, float* out, cudaStream_t s) {
  const int64_t total = (int64_t)B * T * C;
  if (total == 0) return;
  // scale = (do_scale ? sqrt(C) : 1); the Conformer's positional-encoding front end multiplies by
  // sqrt(attention_dim) again (esp/transformer/embedding.py:253) -- applied as a second factor.
  embedding_kernel<<<blocks_for(total, 256), 256, 0, s>>>(ids, len, table, T, C, vocab, scale, sqrtf((float)C), out,
                                                         total);
  PT_LAUNCHED();
}

void glu_dw_bn_swish_cl(const float* in, const int64_t* len, const float* dw_w, const float* dw_b,
                        const float* bn_scale, const float* bn_shift, int B, int T, int C, int K, float* out,
                        cudaStream_t s) {
  const int64_t total = (int64_t)B * T * C;
  if (total == 0) return;
  glu_dw_bn_swish_kernel<<<blocks_for(total, 256), 256, 0, s>>>(in, len, dw_w, dw_b, bn_scale, bn_shift, T, C, K, out,
                                                               total);
  PT_LAUNCHED();
}

void l2_normalize_rows(float* x, int rows, int C, cudaStream_t s) {
  if (rows == 0) return;
  l2norm_rows_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, rows, C);
  PT_LAUNCHED();
}

void style_mdn_sample(const float* logpi, const float* logsigma, const float* mu, const float* z, int B, int G, int D,
                      float noise_scale, int normalize, float* style, cudaStream_t s, const float* comp_u) {
  if (B == 0) return;
  style_mdn_sample_kernel<<<B, 256, 0, s>>>(logpi, logsigma, mu, z, G, D, noise_scale, normalize, style, comp_u);
  PT_LAUNCHED();
}

void add_row_broadcast(float* x, const float* v, int B, int T, int C, cudaStream_t s) {
  const int64_t total = (int64_t)B * T * C;
  if (total == 0) return;
  add_row_broadcast_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, v, T, C, total);
  PT_LAUNCHED();
}

void mdn_duration_head(const float* h, const float* w_pi, const float* b_pi, const float* w_ls, const float* b_ls,
                       const float* w_mu, const float* b_mu, int rows, int C, int G, float* log_d, cudaStream_t s) {
  PT_CHECK(G <= MAX_G, "mdn_duration_head: num_gaussians=%d > %d", G, MAX_G);
  if (rows == 0) return;
  mdn_duration_head_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(h, w_pi, b_pi, w_ls, b_ls, w_mu, b_mu, rows, C, G, log_d);
  PT_LAUNCHED();
}

void duration_quantize(const float* log_d, const int64_t* phone_len, int B, int Tx, int64_t* dur, int64_t* frame_len,
                       cudaStream_t s) {
  PT_CHECK(log_d && dur && frame_len, "duration_quantize: null pointer");
  if (B == 0) return;
  duration_quantize_kernel<<<B, 256, 0, s>>>(log_d, phone_len, Tx, dur, frame_len);
  PT_LAUNCHED();
}

void length_regulate(const float* x, const int64_t* dur, int B, int Tx, int C, int Ty, float* out, int32_t* idx_out,
                     cudaStream_t s) {
  PT_CHECK(x && dur && out, "length_regulate: null pointer");
  PT_CHECK(C % 4 == 0 && aligned16(x) && aligned16(out), "length_regulate: C %% 4 == 0 and 16-byte alignment required");
  PT_CHECK(Tx * sizeof(int) <= 200 * 1024, "length_regulate: Tx=%d too long", Tx);
  if (B == 0 || Ty == 0) return;
  const size_t smem = (size_t)Tx * sizeof(int);
  if (smem > 48 * 1024)
    PT_CUDA(cudaFuncSetAttribute(length_regulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(Ty, LR_ROWS), B);
  length_regulate_kernel<<<grid, 256, smem, s>>>(x, dur, Tx, C, Ty, out, idx_out);
  PT_LAUNCHED();
}

void pitch_head(const float* h, const float* w, const float* b, const int64_t* len, int B, int T, int C,
                float* log_cf0, float* vuv, cudaStream_t s) {
  const int64_t rows = (int64_t)B * T;
  if (rows == 0) return;
  pitch_head_kernel<<<blocks_for(rows, 8), 256, 0, s>>>(h, w, b, len, B, T, C, log_cf0, vuv);
  PT_LAUNCHED();
}

void pitch_embed_add(float* x, const float* log_cf0, const float* w, const float* b, const int64_t* len, int B, int T,
                     int C, cudaStream_t s) {
  const int64_t total = (int64_t)B * T * C;
  if (total == 0) return;
  pitch_embed_add_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, log_cf0, w, b, len, T, C, total);
  PT_LAUNCHED();
}

void ddpm_update(float* x, const float* eps, const float* z, int B, int T, int M, float c_recip, float c_recipm1,
                 float coef1, float coef2, float sigma, cudaStream_t s, void* xp_hi, void* xp_lo, int Mp) {
  if (B == 0 || T == 0) return;
  dim3 grid(ceil_div(T, 32), ceil_div(M, 32), B);
  ddpm_update_kernel<<<grid, 256, 0, s>>>(x, eps, z, T, M, c_recip, c_recipm1, coef1, coef2, sigma, (__half*)xp_hi,
                                          (__half*)xp_lo, Mp);
  PT_LAUNCHED();
}

}  // namespace pttspp

extern "C" int pttspp_duration_quantize(const float* log_d, const int64_t* phone_len, int B, int Tx, int64_t* dur,
                                        int64_t* frame_len, pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::duration_quantize(log_d, phone_len, B, Tx, dur, frame_len, (cudaStream_t)stream);
  PT_API_END
}

extern "C" int pttspp_length_regulate(const float* x, const int64_t* dur, int B, int Tx, int C, int Ty, float* out,
                                      int32_t* idx_out, pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::length_regulate(x, dur, B, Tx, C, Ty, out, idx_out, (cudaStream_t)stream);
  PT_API_END
}

namespace pttspp {
void iir_filtfilt(const float* x, float* y, float* tmp, int rows, int T, const float* b, const float* a, int ntaps,
                  const int64_t* row_len, int min_len,
                  cudaStream_t s) {
  PT_CHECK(x && y && tmp && b && a, "iir_filtfilt: null pointer");
  PT_CHECK(ntaps >= 1 && ntaps <= IIR_MAX_TAPS, "iir_filtfilt: 1..%d coefficients supported", IIR_MAX_TAPS);
  if (rows <= 0 || T <= 0) return;
  iir_filtfilt_kernel<<<ceil_div(rows, 64), 64, 0, s>>>(x, y, tmp, rows, T, b, a, ntaps, row_len, min_len);
  PT_LAUNCHED();
}
}  // namespace pttspp

extern "C" int pttspp_iir_filtfilt(const float* x, float* y, float* scratch, int rows, int T, const float* b_coeffs,
                                   const float* a_coeffs, int ntaps, pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::iir_filtfilt(x, y, scratch, rows, T, b_coeffs, a_coeffs, ntaps, nullptr, 0, (cudaStream_t)stream);
  PT_API_END
}

extern "C" int pttspp_iir_filtfilt_ragged(const float* x, float* y, float* scratch, int rows, int T, const int64_t* row_len,
                                          int min_len, const float* b_coeffs, const float* a_coeffs, int ntaps,
                                          pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(row_len != nullptr && min_len >= 0, "iir_filtfilt_ragged: row_len is required");
  pttspp::iir_filtfilt(x, y, scratch, rows, T, b_coeffs, a_coeffs, ntaps, row_len, min_len, (cudaStream_t)stream);
  PT_API_END
}
