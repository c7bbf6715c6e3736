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
#include "aa_math.cuh"

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
    // sin(alpha*u): the fast sine (x * 1/2pi rounded toward zero, then the SFU's sine of the fractional revolution) --
    // |error| ~ ulp(alpha*u / 2pi) * 2pi < 1e-5 over the activations' range, far inside the 1e-4 RMS waveform bar; libm
    // sinf's slow path made this kernel instruction bound, and an explicit r - rint(r) reduction costs two more FP32-pipe
    // operations plus an FRND on the quarter-rate XU pipe for the same argument error (ncu: FMA pipe 60 %, XU 52 %)
    const float sn = __sinf(u * a2);
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
        const __half h = pt_f2h_sat(acc);
        hp[(int64_t)t * C] = h;
        lp[(int64_t)t * C] = pt_f2h_sat(acc - __half2float(h));
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
  const float a2 = alpha;
  const float* xb = x + (int64_t)b * L * C + c;
  const int64_t ob = (int64_t)b * L * C + c + (int64_t)t0 * C;
  // warp-uniform: a warp spans 32 channels of ONE strip
  if (t0 >= 5 && t0 + TT + 5 <= L) aa_snake_strip<false>(xb, y, y_hi, y_lo, ob, t0, L, C, f2, g, a2, inv_alpha);
  else aa_snake_strip<true>(xb, y, y_hi, y_lo, ob, t0, L, C, f2, g, a2, inv_alpha);
}


// ---- channel-pair kernel (the production path) ----------------------------------------------------------------------
// Same arithmetic, two adjacent channels per thread in packed fp32x2 registers (Blackwell's FFMA2 / FMUL2 / FADD2 halve the
// instruction count of this instruction-bound kernel) and a ROLLING strip: a thread walks AP_NB blocks of AP_TB outputs,
// carrying the 10 up-sampled values and 5 inputs that neighbouring blocks share instead of recomputing / reloading the
// halo (the strip kernel above re-reads 26 inputs and recomputes 42 up-sampled values per 16 outputs: 1.6x read
// amplification).  Needs exactly symmetric filters, f[k] == f[11-k] (true for the reference's Kaiser sinc: checked on the
// host where the filters are loaded): 6 + 6 packed coefficients stay in registers.  With symmetric filters every
// channel's result is bit-identical to the strip kernel's (same operation order).
constexpr int AP_STRIP = 64;  // outputs per thread
constexpr int AP_TB = 8;      // outputs per block; a strip is AP_STRIP / AP_TB = 8 blocks

// The kernel is ISSUE bound (timing experiments in profiles/r02c_aa_experiments.txt: removing the sine, the stores or the
// down-sampling taps each shortens it in proportion to the removed instructions; an FFMA2 holds the issue port for two
// cycles, so 37 packed + 37 other instructions per output pair are 111 of the 129 measured cycles).  What can go is the
// "other" half: the rolling windows therefore live in RING buffers with compile-time slot numbers -- 32 slots for the
// up-sampled values (window 26, advance 16 per block), 16 for the inputs (window 13, advance 8): after two blocks every
// value is back in its slot, so the loop body is two blocks and the 15 packed register moves per block that carried the
// halo are gone -- and the output mode (fp32 / operand planes) is a template parameter instead of a branch per store.
// P = block parity.  S(i): ring slot of window entry i (entry i <-> up-sampled index 2*tb - 5 + i); X(i): slot of input
// window entry i (entry i <-> x[tb - 5 + i] ... the five carried inputs are entries 0..4, the block's new ones 5..12).
template <bool EDGE, int P, int OM>
__device__ __forceinline__ void aa_ring_block(const float* __restrict__ xb, float* __restrict__ y, __half* __restrict__ y_hi,
                                              __half* __restrict__ y_lo, const int64_t ob, const int tb, const int L,
                                              const int ldc, f32x2 (&sv)[32], f32x2 (&xw)[16], const bool more,
                                              const f32x2 (&e)[6], const f32x2 (&g)[6], const f32x2 a2,
                                              const f32x2 inv_alpha) {
#define AA_S(i) ((16 * P + (i)) & 31)
#define AA_X(i) ((8 * P + (i)) & 15)
  // new up-sampled values: entries 10 .. 25
#pragma unroll
  for (int i = 10; i < 2 * AP_TB + 10; ++i) {
    f32x2 u = pk2(0.f, 0.f);
    if ((i & 1) == 0) {
#pragma unroll
      for (int dd = 0; dd < 6; ++dd) u = fma2(xw[AA_X(i / 2 + dd - 5)], e[dd], u);          // f2[10 - 2d]
    } else {
#pragma unroll
      for (int dd = 0; dd < 6; ++dd) u = fma2(xw[AA_X((i - 1) / 2 + dd - 5)], e[5 - dd], u);  // f2[11 - 2d] == f2[2d]
    }
    sv[AA_S(i)] = snake2(u, a2, inv_alpha);
  }
  // the next block's inputs x[tb + 13 .. tb + 20] go into the slots of input entries 0..7 of this block, which the
  // up-sampling above has finished with (entries 8..12 are the next block's carried inputs); they have the whole
  // down-sampling phase to arrive
  if (more) {
    if (tb + 2 * AP_TB + 4 <= L - 1) {  // rows inside the utterance: one pointer, compile-time multiples of the row stride
      const float* xp = xb + (int64_t)(tb + AP_TB + 5) * ldc;
#pragma unroll
      for (int i = 0; i < AP_TB; ++i) xw[AA_X(13 + i)] = *reinterpret_cast<const f32x2*>(xp + i * ldc);
    } else {
#pragma unroll
      for (int i = 0; i < AP_TB; ++i) {
        int l = tb + AP_TB + 5 + i;
        l = l > L - 1 ? L - 1 : l;
        xw[AA_X(13 + i)] = *reinterpret_cast<const f32x2*>(xb + (int64_t)l * ldc);
      }
    }
  }
  if (EDGE) {
    const int imax = 2 * (L - tb) + 4;  // entry of up-sampled sample 2L-1: later ones repeat it
#pragma unroll
    for (int i = 10; i < 2 * AP_TB + 10; ++i)
      if (i > imax) sv[AA_S(i)] = sv[AA_S(i - 1)];
  }
  float* const yb = (OM & 1) ? y + ob : nullptr;  // block bases: every store is base + (compile-time t) * row stride
  __half* const hb = (OM & 2) ? y_hi + ob : nullptr;
  __half* const lb = (OM & 2) ? y_lo + ob : nullptr;
#pragma unroll
  for (int t = 0; t < AP_TB; ++t) {
    if (!EDGE || tb + t < L) {
      f32x2 acc = pk2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 12; ++k) acc = fma2(sv[AA_S(2 * t + k)], g[k < 6 ? k : 11 - k], acc);
      float a0, a1;
      upk2(acc, a0, a1);
      if (OM & 1) *reinterpret_cast<float2*>(yb + t * ldc) = make_float2(a0, a1);
      if (OM & 2) {
        const __half2 h = pt_f2h2_sat(a0, a1);
        const float2 hf = __half22float2(h);
        *reinterpret_cast<__half2*>(hb + t * ldc) = h;
        *reinterpret_cast<__half2*>(lb + t * ldc) = pt_f2h2_sat(a0 - hf.x, a1 - hf.y);
      }
    }
  }
#undef AA_S
#undef AA_X
}

template <int OM>
__global__ void __launch_bounds__(256, 2) aa_snake_pair_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               __half* __restrict__ y_hi, __half* __restrict__ y_lo,
                                                               int L, int C, int n_strips,
                                                               const float* __restrict__ log_alpha,
                                                               const float* __restrict__ up_f,
                                                               const float* __restrict__ down_f) {
  const int half_c = C >> 1;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cp = (int)(gid % half_c);
  const long long strip_ll = gid / half_c;  // blockIdx.y = utterance
  if (strip_ll >= n_strips) return;
  const int strip = (int)strip_ll;
  const int c = 2 * cp;
  const int t0 = strip * AP_STRIP;
  const int ldc = C;
  f32x2 e[6], g[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    const float ef = 2.f * up_f[10 - 2 * d];
    e[d] = pk2(ef, ef);
    g[d] = pk2(down_f[d], down_f[d]);
  }
  const float al0 = expf(log_alpha[c]), al1 = expf(log_alpha[c + 1]);
  const f32x2 inv_alpha = pk2(1.f / (al0 + 1e-9f), 1.f / (al1 + 1e-9f));
  const f32x2 a2 = pk2(al0, al1);
  const float* xb = x + (int64_t)blockIdx.y * L * C + c;
  f32x2 sv[32], xw[16];
  // prologue: the 10 up-sampled values in front of the strip (indices 2*t0 - 5 .. 2*t0 + 4) from x[t0 - 5 .. t0 + 4]
  // -> ring entries 0..9 of block parity 0; inputs x[t0 .. t0 + 4] -> input entries 0..4, x[t0 + 5 .. t0 + 12] -> 5..12
  {
    f32x2 xs[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      int l = t0 - 5 + i;
      l = l < 0 ? 0 : (l > L - 1 ? L - 1 : l);
      xs[i] = *reinterpret_cast<const f32x2*>(xb + (int64_t)l * ldc);
    }
#pragma unroll
    for (int i = 0; i < AP_TB; ++i) {
      int l = t0 + 5 + i;
      l = l > L - 1 ? L - 1 : l;
      xw[5 + i] = *reinterpret_cast<const f32x2*>(xb + (int64_t)l * ldc);
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      f32x2 u = pk2(0.f, 0.f);
      if ((i & 1) == 0) {
#pragma unroll
        for (int dd = 0; dd < 6; ++dd) u = fma2(xs[i / 2 + dd], e[dd], u);
      } else {
#pragma unroll
        for (int dd = 0; dd < 6; ++dd) u = fma2(xs[(i - 1) / 2 + dd], e[5 - dd], u);
      }
      sv[i] = snake2(u, a2, inv_alpha);
    }
    // replicate padding of the up-sampled signal at both ends of the utterance
#pragma unroll
    for (int i = 4; i >= 0; --i)
      if (2 * t0 - 5 + i < 0) sv[i] = sv[i + 1];
    const int imax = 2 * (L - t0) + 4;
#pragma unroll
    for (int i = 1; i < 10; ++i)
      if (i > imax) sv[i] = sv[i - 1];
#pragma unroll
    for (int i = 0; i < 5; ++i) xw[i] = xs[5 + i];
  }
  const int64_t ob0 = (int64_t)blockIdx.y * L * C + c;
  constexpr int AP_NB = AP_STRIP / AP_TB;
#pragma unroll 1
  for (int blk = 0; blk < AP_NB; blk += 2) {
    {
      const int tb = t0 + blk * AP_TB;
      if (tb >= L) break;
      const int64_t ob = ob0 + (int64_t)tb * ldc;
      const bool more = tb + AP_TB < L;  // (blk + 1 < AP_NB always: blk is even)
      if (tb + AP_TB + 4 <= L - 1) aa_ring_block<false, 0, OM>(xb, y, y_hi, y_lo, ob, tb, L, ldc, sv, xw, more, e, g, a2, inv_alpha);
      else aa_ring_block<true, 0, OM>(xb, y, y_hi, y_lo, ob, tb, L, ldc, sv, xw, more, e, g, a2, inv_alpha);
    }
    {
      const int tb = t0 + (blk + 1) * AP_TB;
      if (tb >= L) break;
      const int64_t ob = ob0 + (int64_t)tb * ldc;
      const bool more = (blk + 2 < AP_NB) && (tb + AP_TB < L);
      if (tb + AP_TB + 4 <= L - 1) aa_ring_block<false, 1, OM>(xb, y, y_hi, y_lo, ob, tb, L, ldc, sv, xw, more, e, g, a2, inv_alpha);
      else aa_ring_block<true, 1, OM>(xb, y, y_hi, y_lo, ob, tb, L, ldc, sv, xw, more, e, g, a2, inv_alpha);
    }
  }
}

}  // namespace

void aa_snake_cl(const float* x, float* y, int B, int L, int C, const float* log_alpha, const float* up_f,
                 const float* down_f, cudaStream_t s, void* y_hi, void* y_lo, int symmetric_filters) {
  PT_CHECK(x && (y || y_hi) && log_alpha && up_f && down_f, "aa_snake: null pointer");
  PT_CHECK(!y_hi || y_lo, "aa_snake: y_hi without y_lo");
  PT_CHECK(x != y, "aa_snake: in-place operation is not supported");
  PT_CHECK(B >= 1 && B <= 65535 && L >= 1 && C >= 1, "aa_snake: bad shape");
  ProfScope prof(PROF_AA_SNAKE, s, 0.0, 2.0 * 4.0 * B * (double)L * C);
  auto a8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7) == 0; };
  if (symmetric_filters && C % 2 == 0 && a8(x) && (!y || a8(y)) && (!y_hi || (a8(y_hi) && a8(y_lo)))) {
    const int n_strips = ceil_div(L, AP_STRIP);
    const long long threads = (long long)(C / 2) * n_strips;
    dim3 grid((unsigned)ceil_div64(threads, 256), B);
    if (y && y_hi)
      aa_snake_pair_kernel<3><<<grid, 256, 0, s>>>(x, y, (__half*)y_hi, (__half*)y_lo, L, C, n_strips, log_alpha, up_f, down_f);
    else if (y_hi)
      aa_snake_pair_kernel<2><<<grid, 256, 0, s>>>(x, y, (__half*)y_hi, (__half*)y_lo, L, C, n_strips, log_alpha, up_f, down_f);
    else
      aa_snake_pair_kernel<1><<<grid, 256, 0, s>>>(x, y, (__half*)y_hi, (__half*)y_lo, L, C, n_strips, log_alpha, up_f, down_f);
    PT_LAUNCHED();
    return;
  }
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

extern "C" int pttspp_aa_snake_pair_cl(const float* x, float* y, int B, int L, int C, const float* log_alpha,
                                       const float* up_filter, const float* down_filter, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(C % 2 == 0, "aa_snake_pair: the channel-pair kernel needs an even channel count (got %d)", C);
  pttspp::aa_snake_cl(x, y, B, L, C, log_alpha, up_filter, down_filter, (cudaStream_t)stream, nullptr, nullptr, 1);
  PT_API_END
}
