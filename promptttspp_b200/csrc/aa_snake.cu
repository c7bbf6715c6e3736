// Fused anti-aliased Snake activation on channels-last activations x[b][l][c]:
//   u = 2x polyphase up-sample of x (replicate pad 5, 12-tap Kaiser sinc, gain 2, crop 15/15)
//   s = u + sin^2(alpha*u) / (alpha + 1e-9),  alpha = exp(log_alpha[c])
//   y = 2x low-pass down-sample of s (replicate pad 5 left / 6 right, 12 taps, stride 2)
// Closed forms (derived from promptttspp/layers/activations.py:74-138, checked by the oracle):
//   u[2j]   = 2 * sum_{d<6} x[clamp(j-3+d)] * f[11-2d]
//   u[2j+1] = 2 * sum_{d<6} x[clamp(j-2+d)] * f[10-2d]
//   y[t]    = sum_{k<12} s[clamp(2t+k-5, 0, 2L-1)] * g[k]
// A thread owns one channel and a strip of TT consecutive outputs: 2*TT+10 up-sampled values are
// produced in registers and never touch memory.  A warp spans 32 consecutive channels, so every
// row access is one coalesced 128-byte line.  HBM-bound: 2*4 bytes per element.
#include <cuda_fp16.h>

#include "common.h"

namespace pttspp {
namespace {

constexpr int TT = 16;          // outputs per thread
constexpr int NX = TT + 10;     // input samples per strip
constexpr int NS = 2 * TT + 10; // up-sampled samples per strip

// y (fp32) and/or split-fp16 planes y_hi/y_lo (operands of the tcgen05 convs that consume the activation).
// f2 = 2 * up filter (the x2 gain folded in: exact).  EDGE = false: strips whose halo lies inside [0, L) -- no index
// clamps, no replicate fix-ups (the kernel is instruction bound, ncu: SM 79 % busy at 1.8 TB/s, so every removed
// instruction is time); EDGE = true: the general path for the first / last strips of an utterance.
template <bool EDGE>
__device__ __forceinline__ void aa_snake_strip(const float* __restrict__ xb, float* __restrict__ y,
                                               __half* __restrict__ y_hi, __half* __restrict__ y_lo, int64_t ob, int t0,
                                               int L, int C, const float (&f2)[12], const float (&g)[12], float a2,
                                               float inv_alpha) {
  float xs[NX];
  if (EDGE) {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      int l = t0 - 5 + i;
      l = l < 0 ? 0 : (l > L - 1 ? L - 1 : l);
      xs[i] = xb[(int64_t)l * C];
    }
  } else {
    const float* xp = xb + (int64_t)(t0 - 5) * C;
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = xp[(int64_t)i * C];
  }
  // up-sample + snake.  s[i] <-> up-sampled index m = 2*t0 - 5 + i:
  //   i even -> m odd,  j = t0 - 3 + i/2, taps x[j-2+d] = xs[i/2 + d],      filter f[10-2d]
  //   i odd  -> m even, j = t0 - 2 + (i-1)/2, taps x[j-3+d] = xs[(i-1)/2+d], filter f[11-2d]
  float sv[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    float u = 0.f;
    if ((i & 1) == 0) {
#pragma unroll
      for (int dd = 0; dd < 6; ++dd) u = fmaf(xs[i / 2 + dd], f2[10 - 2 * dd], u);
    } else {
#pragma unroll
      for (int dd = 0; dd < 6; ++dd) u = fmaf(xs[(i - 1) / 2 + dd], f2[11 - 2 * dd], u);
    }
    // sin(alpha*u) with the argument reduced in REVOLUTIONS: r = u * alpha/(2*pi), frac = r - rint(r) is exact, and
    // the MUFU sine of 2*pi*frac in [-pi, pi] is accurate to 2^-21 -- |error| < 1e-6 over the activations' range, far
    // inside the 1e-4 RMS waveform bar (libm sinf's slow path made this kernel instruction bound)
    const float r = u * a2;
    const float frac = r - rintf(r);
    const float sn = __sinf(frac * 6.28318530717958648f);
    sv[i] = fmaf(inv_alpha * sn, sn, u);
  }
  if (EDGE) {
    // replicate padding of the up-sampled signal: indices below 0 / above 2L-1 repeat the edge value
#pragma unroll
    for (int i = 4; i >= 0; --i)
      if (2 * t0 - 5 + i < 0) sv[i] = sv[i + 1];
    const int imax = 2 * (L - t0) + 4;  // strip index of up-sampled sample 2L-1
#pragma unroll
    for (int i = 1; i < NS; ++i)
      if (i > imax) sv[i] = sv[i - 1];
  }
  float* yp = y ? y + ob : nullptr;
  __half* hp = y_hi ? y_hi + ob : nullptr;
  __half* lp = y_hi ? y_lo + ob : nullptr;
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    if (!EDGE || t0 + t < L) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 12; ++k) acc = fmaf(sv[2 * t + k], g[k], acc);
      if (yp) yp[(int64_t)t * C] = acc;
      if (hp) {
        const __half h = __float2half_rn(acc);
        hp[(int64_t)t * C] = h;
        lp[(int64_t)t * C] = __float2half_rn(acc - __half2float(h));
      }
    }
  }
}

__global__ void __launch_bounds__(256) aa_snake_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       __half* __restrict__ y_hi, __half* __restrict__ y_lo, int L,
                                                       int C, const float* __restrict__ log_alpha,
                                                       const float* __restrict__ up_f,
                                                       const float* __restrict__ down_f) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int t0 = (blockIdx.y * blockDim.y + threadIdx.y) * TT;
  const int b = blockIdx.z;
  if (c >= C || t0 >= L) return;
  float f2[12], g[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    f2[i] = 2.f * up_f[i];
    g[i] = down_f[i];
  }
  const float alpha = expf(log_alpha[c]);
  const float inv_alpha = 1.f / (alpha + 1e-9f);
  const float a2 = alpha * 0.15915494309189535f;
  const float* xb = x + (int64_t)b * L * C + c;
  const int64_t ob = (int64_t)b * L * C + c + (int64_t)t0 * C;
  // warp-uniform: a warp spans 32 channels of ONE strip
  if (t0 >= 5 && t0 + TT + 5 <= L) aa_snake_strip<false>(xb, y, y_hi, y_lo, ob, t0, L, C, f2, g, a2, inv_alpha);
  else aa_snake_strip<true>(xb, y, y_hi, y_lo, ob, t0, L, C, f2, g, a2, inv_alpha);
}

}  // namespace

void aa_snake_cl(const float* x, float* y, int B, int L, int C, const float* log_alpha, const float* up_f,
                 const float* down_f, cudaStream_t s, void* y_hi, void* y_lo) {
  PT_CHECK(x && (y || y_hi) && log_alpha && up_f && down_f, "aa_snake: null pointer");
  PT_CHECK(!y_hi || y_lo, "aa_snake: y_hi without y_lo");
  PT_CHECK(x != y, "aa_snake: in-place operation is not supported");
  PT_CHECK(B >= 1 && B <= 65535 && L >= 1 && C >= 1, "aa_snake: bad shape");
  ProfScope prof(PROF_AA_SNAKE, s, 0.0, 2.0 * 4.0 * B * (double)L * C);
  dim3 block(32, 8);
  dim3 grid(ceil_div(C, 32), ceil_div(L, 8 * TT), B);
  PT_CHECK(grid.y <= 65535, "aa_snake: L=%d too long for one launch", L);
  aa_snake_kernel<<<grid, block, 0, s>>>(x, y, (__half*)y_hi, (__half*)y_lo, L, C, log_alpha, up_f, down_f);
  PT_LAUNCHED();
}

}  // namespace pttspp

extern "C" int pttspp_aa_snake_cl(const float* x, float* y, int B, int L, int C, const float* log_alpha,
                                  const float* up_filter, const float* down_filter, pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::aa_snake_cl(x, y, B, L, C, log_alpha, up_filter, down_filter, (cudaStream_t)stream);
  PT_API_END
}
