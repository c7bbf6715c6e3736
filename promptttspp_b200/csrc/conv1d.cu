// Implicit-GEMM Conv1d on channels-last fp32 activations, CUDA-core (FFMA) path.
//
// One kernel serves every dense contraction of the hot path that is not (yet) routed to the
// tcgen05 path (conv1d_umma.cu): Linear (K=1), dilated / wide Conv1d, polyphase ConvTranspose1d
// (out_mul/out_off), with the fused prologue/epilogue described in include/pttspp_b200.h.
//
// Tiling: a CTA computes BM output rows x BN output channels of one batch item.  Per 16-channel
// slab of Cin it stages the (BM-1)*in_stride + (K-1)*dil + 1 input rows that all K taps touch in
// shared memory once, then streams the K weight slabs [16][BN] through a second buffer; every
// thread owns a TM x TN register tile (rows strided by BM/TM so that the two/four row groups of
// a warp hit different banks).  Full fp32 accumulation: this path is the numerical anchor.
#include "common.h"
#include "conv_epilogue.cuh"

namespace pttspp {

namespace {

constexpr int BK = 16;
constexpr int LDA = BK + 4;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), (BM * BN >= 128 * 128) ? 2 : 2)
conv1d_simt_kernel(const pttspp_conv1d_desc d) {
  constexpr int TXN = BN / TN;  // threads along channels
  constexpr int TYN = BM / TM;  // threads along rows
  constexpr int NT = TXN * TYN;
  constexpr int NG = TN / 4;    // float4 column groups per thread
  constexpr int GS = BN / NG;   // column stride between groups
  extern __shared__ __align__(16) float smem[];

  const int b = blockIdx.z;
  const int m0 = d.m_begin + blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;

  const int rowsA = (BM - 1) * d.in_stride + (d.K - 1) * d.dil + 1;
  float* As = smem;
  float* Bs = smem + rowsA * LDA;

  int len_in = d.T_in;
  if (d.in_len) {
    long long l = d.in_len[b];
    len_in = (int)(l < (long long)d.T_in ? l : (long long)d.T_in);
  }
  const int row_base = m0 * d.in_stride - d.pad;
  const float* in_b = d.in + (int64_t)b * d.in_bs;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < d.Cin; c0 += BK) {
    // ---- stage the input rows of this channel slab (zero outside the valid range) ----
    for (int i = tid; i < rowsA * (BK / 4); i += NT) {
      const int r = i / (BK / 4), q = i % (BK / 4);
      const int gr = row_base + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr >= 0 && gr < len_in) {
        v = *reinterpret_cast<const float4*>(in_b + (int64_t)gr * d.in_ld + c0 + q * 4);
        if (d.in_add) {
          const float4 a = *reinterpret_cast<const float4*>(d.in_add + c0 + q * 4);
          v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
      }
      *reinterpret_cast<float4*>(As + r * LDA + q * 4) = v;
    }
    for (int k = 0; k < d.K; ++k) {
      // ---- weight slab [BK][BN] of tap k ----
      for (int i = tid; i < BK * (BN / 4); i += NT) {
        const int kk = i / (BN / 4), c4 = i % (BN / 4);
        const int col = n0 + c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < d.w_ld)
          v = *reinterpret_cast<const float4*>(d.w + ((int64_t)k * d.Cin + c0 + kk) * d.w_ld + col);
        *reinterpret_cast<float4*>(Bs + kk * BN + c4 * 4) = v;
      }
      __syncthreads();
      const float* Ap = As + (k * d.dil) * LDA;
#pragma unroll
      for (int kk4 = 0; kk4 < BK / 4; ++kk4) {
        float4 a[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i)
          a[i] = *reinterpret_cast<const float4*>(Ap + ((ty + TYN * i) * d.in_stride) * LDA + kk4 * 4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float4 bv[NG];
#pragma unroll
          for (int g = 0; g < NG; ++g)
            bv[g] = *reinterpret_cast<const float4*>(Bs + (kk4 * 4 + e) * BN + g * GS + tx * 4);
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            const float av = (e == 0) ? a[i].x : (e == 1) ? a[i].y : (e == 2) ? a[i].z : a[i].w;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              acc[i][g * 4 + 0] = fmaf(av, bv[g].x, acc[i][g * 4 + 0]);
              acc[i][g * 4 + 1] = fmaf(av, bv[g].y, acc[i][g * 4 + 1]);
              acc[i][g * 4 + 2] = fmaf(av, bv[g].z, acc[i][g * 4 + 2]);
              acc[i][g * 4 + 3] = fmaf(av, bv[g].w, acc[i][g * 4 + 3]);
            }
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- epilogue (conv_epilogue.cuh) ----
  long long olen = 0x7fffffffffffffffLL;
  if (d.out_len) olen = d.out_len[b];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + TYN * i;
    if (m >= d.m_begin + d.M) continue;
    const int row = m * d.out_mul + d.out_off;
    if (row < 0 || row >= d.T_out) continue;
    const float mask = ((long long)row < olen) ? 1.f : 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float a4[4] = {acc[i][g * 4 + 0], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]};
      conv_epilogue4(d, b, row, mask, n0 + g * GS + tx * 4, a4);
    }
  }
}

template <int BM, int BN, int TM, int TN>
void launch_simt(const pttspp_conv1d_desc& d, cudaStream_t s) {
  const int rowsA = (BM - 1) * d.in_stride + (d.K - 1) * d.dil + 1;
  const size_t smem = (size_t)(rowsA * LDA + BK * BN) * sizeof(float);
  PT_CHECK(smem <= 220 * 1024, "conv1d: tile of %d input rows does not fit shared memory", rowsA);
  auto kern = conv1d_simt_kernel<BM, BN, TM, TN>;
  if (smem > 48 * 1024)
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(d.M, BM), ceil_div(d.Cout, BN), d.B);
  kern<<<grid, (BM / TM) * (BN / TN), smem, s>>>(d);
  PT_LAUNCHED();
}

}  // namespace

void conv1d_umma_cl(const pttspp_conv1d_desc& d, cudaStream_t s);  // conv1d_umma.cu
bool conv1d_umma_supported(const pttspp_conv1d_desc& d);

void conv1d_cl(const pttspp_conv1d_desc& d, cudaStream_t s) {
  PT_CHECK(d.out || d.out_hi, "conv1d: no output (out and out_hi are NULL)");
  PT_CHECK(!d.out_hi || d.out_lo, "conv1d: out_hi without out_lo");
  PT_CHECK(!d.res_hi || (d.res_lo && !d.res && (d.impl == 2 || d.impl == 3 || d.impl == 4 || (d.impl == 0 && conv1d_umma_supported(d)))),
           "conv1d: res_hi needs res_lo, no fp32 res, and the tcgen05 path");
  PT_CHECK(d.Cin > 0 && d.Cin % BK == 0, "conv1d: Cin=%d must be a positive multiple of %d", d.Cin, BK);
  PT_CHECK(!d.in_add || aligned16(d.in_add), "conv1d: in_add must be 16-byte aligned");
  PT_CHECK(d.K >= 1 && d.dil >= 1 && d.in_stride >= 1 && d.out_mul >= 1, "conv1d: bad geometry");
  PT_CHECK(d.act != PTTSPP_ACT_GATE || d.Cout % 2 == 0, "conv1d: gate activation needs even Cout");
  PT_CHECK(d.B >= 1 && d.B <= 65535, "conv1d: batch %d out of range", d.B);
  if (d.M <= 0 || d.Cout <= 0) return;
  const double flops = 2.0 * d.B * (double)d.M * d.Cout * (double)d.Cin * d.K;
  const bool umma = (d.impl == 2) || (d.impl == 3) || (d.impl == 4) || (d.impl == 0 && conv1d_umma_supported(d));
  ProfScope prof(umma ? PROF_CONV_UMMA : PROF_CONV_SIMT, s, flops, 0.0);
  if (d.impl == 2 || d.impl == 3 || d.impl == 4) {
    PT_CHECK(conv1d_umma_supported(d), "conv1d: tcgen05 path requested for an unsupported shape");
    conv1d_umma_cl(d, s);
    return;
  }
  if (d.impl == 0 && conv1d_umma_supported(d)) {
    conv1d_umma_cl(d, s);
    return;
  }
  PT_CHECK(d.in && d.w, "conv1d: the CUDA-core path needs fp32 `in` and packed fp32 weights `w`");
  PT_CHECK(d.in_ld % 4 == 0 && aligned16(d.in), "conv1d: input must be 16-byte aligned, ld %% 4 == 0");
  PT_CHECK(d.w_ld % 4 == 0 && d.w_ld >= d.Cout && aligned16(d.w), "conv1d: packed weight ld=%d invalid", d.w_ld);
  // few large tiles leave SMs idle on short sequences (the encoder's 1024 -> 256 k9 conv at Tx = 256 ran 64 CTAs):
  // take the 64-column tile when the 128-column grid would not fill the machine
  const long long ctas128 = (long long)ceil_div(d.M, 128) * ceil_div(d.Cout, 128) * d.B;
  if (d.Cout > 64 && ctas128 >= 120)
    launch_simt<128, 128, 8, 8>(d, s);
  else if (d.Cout > 32)
    launch_simt<128, 64, 8, 4>(d, s);
  else
    launch_simt<256, 32, 8, 4>(d, s);
}

// ---- host-side weight packing (runs once at load; plain C++) -------------------------------

void pack_conv_weight(const float* v, const float* g, int Cout, int Cin, int K, float* packed, int w_ld,
                      int interleave_halves, cudaStream_t) {
  PT_CHECK(w_ld >= Cout && w_ld % 4 == 0, "pack_conv_weight: bad w_ld");
  PT_CHECK(!interleave_halves || Cout % 2 == 0, "pack_conv_weight: interleave needs even Cout");
  std::fill(packed, packed + (size_t)K * Cin * w_ld, 0.f);
  const int half = Cout / 2;
  for (int co = 0; co < Cout; ++co) {
    const float* vr = v + (size_t)co * Cin * K;
    float scale = 1.f;
    if (g) {
      double ss = 0.0;
      for (int i = 0; i < Cin * K; ++i) ss += (double)vr[i] * vr[i];
      scale = (float)((double)g[co] / std::sqrt(ss));
    }
    const int col = interleave_halves ? (co < half ? 2 * co : 2 * (co - half) + 1) : co;
    for (int ci = 0; ci < Cin; ++ci)
      for (int k = 0; k < K; ++k) {
        const float w = vr[(size_t)ci * K + k];
        packed[((size_t)k * Cin + ci) * w_ld + col] = g ? w * scale : w;
      }
  }
}

void pack_convtr_weight(const float* v, const float* g, int Cin, int Cout, int Kt, int stride, float* packed,
                        int w_ld, cudaStream_t) {
  PT_CHECK(w_ld >= Cout && w_ld % 4 == 0, "pack_convtr_weight: bad w_ld");
  PT_CHECK(Kt % stride == 0, "pack_convtr_weight: kernel %d must be a multiple of stride %d", Kt, stride);
  const int J = Kt / stride;
  std::fill(packed, packed + (size_t)stride * J * Cin * w_ld, 0.f);
  for (int ci = 0; ci < Cin; ++ci) {
    const float* vr = v + (size_t)ci * Cout * Kt;
    float scale = 1.f;
    if (g) {
      double ss = 0.0;
      for (int i = 0; i < Cout * Kt; ++i) ss += (double)vr[i] * vr[i];
      scale = (float)((double)g[ci] / std::sqrt(ss));
    }
    for (int r = 0; r < stride; ++r)
      for (int kp = 0; kp < J; ++kp) {
        const int kt = r + (J - 1 - kp) * stride;
        float* dst = packed + (((size_t)r * J + kp) * Cin + ci) * w_ld;
        for (int co = 0; co < Cout; ++co) {
          const float w = vr[(size_t)co * Kt + kt];
          dst[co] = g ? w * scale : w;
        }
      }
  }
}

void pack_conv_weight_split(const float* v, const float* g, int Cout, int Cin, int K, void* w_hi, void* w_lo,
                            int interleave_halves, float* scale_inv) {
  PT_CHECK(!interleave_halves || Cout % 2 == 0, "pack_conv_weight_split: interleave needs even Cout");
  std::vector<float> w((size_t)Cout * Cin * K);
  float mx = 0.f;
  for (int co = 0; co < Cout; ++co) {
    const float* vr = v + (size_t)co * Cin * K;
    float scale = 1.f;
    if (g) {
      double ss = 0.0;
      for (int i = 0; i < Cin * K; ++i) ss += (double)vr[i] * vr[i];
      scale = (float)((double)g[co] / std::sqrt(ss));
    }
    for (int i = 0; i < Cin * K; ++i) {
      const float x = g ? vr[i] * scale : vr[i];
      w[(size_t)co * Cin * K + i] = x;
      mx = std::max(mx, std::fabs(x));
    }
  }
  int e = 0;  // largest power of two with mx * 2^e <= 16384 (keeps hi and lo in fp16's normal range)
  if (mx > 0.f) {
    e = (int)std::floor(std::log2(16384.0 / (double)mx));
    e = std::max(-14, std::min(e, 24));
  }
  const float sc = std::ldexp(1.f, e);
  *scale_inv = std::ldexp(1.f, -e);
  __half* hi = reinterpret_cast<__half*>(w_hi);
  __half* lo = reinterpret_cast<__half*>(w_lo);
  const int half = Cout / 2;
  for (int co = 0; co < Cout; ++co) {
    const int col = interleave_halves ? (co < half ? 2 * co : 2 * (co - half) + 1) : co;
    for (int ci = 0; ci < Cin; ++ci)
      for (int k = 0; k < K; ++k) {
        const float x = w[((size_t)co * Cin + ci) * K + k] * sc;
        const __half h = __float2half_rn(x);
        const size_t idx = ((size_t)k * Cout + col) * Cin + ci;
        hi[idx] = h;
        lo[idx] = __float2half_rn(x - __half2float(h));
      }
  }
}

namespace {
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                        int64_t n, int C, __half* __restrict__ hi,
                                                        __half* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (add) v += add[i % C];
  __half h, l;
  split_f16(v, h, l);
  hi[i] = h;
  lo[i] = l;
}
}  // namespace

namespace {
// rows [B][T][C] -> planes, rows at or beyond len[b] written as zeros (the `x * mask` in front of a conv); 4 elements
// per thread, 16-byte loads / 8-byte stores
__global__ void __launch_bounds__(256) split_f16_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ len,
                                                             int T, int C4, int64_t n4, uint2* __restrict__ hi,
                                                             uint2* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int64_t r = i / C4;
  const int b = (int)(r / T), t = (int)(r - (int64_t)b * T);
  float4 v = reinterpret_cast<const float4*>(x)[i];
  if (len && (int64_t)t >= len[b]) v = make_float4(0.f, 0.f, 0.f, 0.f);
  __half h[4], l[4];
  split_f16(v.x, h[0], l[0]);
  split_f16(v.y, h[1], l[1]);
  split_f16(v.z, h[2], l[2]);
  split_f16(v.w, h[3], l[3]);
  hi[i] = make_uint2((uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
                     (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
  lo[i] = make_uint2((uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
                     (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) split_f16_pad_kernel(const float* __restrict__ x, int64_t rows, int C, int Cp,
                                                            __half* __restrict__ hi, __half* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Cp) return;
  const int64_t r = i / Cp;
  const int c = (int)(i - r * Cp);
  const float v = (c < C) ? x[r * C + c] : 0.f;
  __half h, l;
  split_f16(v, h, l);
  hi[i] = h;
  lo[i] = l;
}
}  // namespace

void split_f16_pad(const float* x, int64_t rows, int C, int Cp, void* hi, void* lo, cudaStream_t s) {
  PT_CHECK(x && hi && lo && C >= 1 && Cp >= C, "split_f16_pad: bad argument");
  if (rows == 0) return;
  ProfScope prof(PROF_OTHER, s, 0.0, 4.0 * (double)rows * (C + Cp));
  split_f16_pad_kernel<<<(unsigned)ceil_div64(rows * Cp, 256), 256, 0, s>>>(x, rows, C, Cp, (__half*)hi, (__half*)lo);
  PT_LAUNCHED();
}

void split_f16_rows(const float* x, int B, int T, int C, const int64_t* len, void* hi, void* lo, cudaStream_t s) {
  PT_CHECK(x && hi && lo && C % 4 == 0 && aligned16(x) && aligned16(hi) && aligned16(lo), "split_f16_rows: bad argument");
  const int64_t n4 = (int64_t)B * T * (C / 4);
  if (n4 == 0) return;
  ProfScope prof(PROF_OTHER, s, 0.0, 8.0 * 4.0 * (double)n4);
  split_f16_rows_kernel<<<(unsigned)ceil_div64(n4, 256), 256, 0, s>>>(x, len, T, C / 4, n4, (uint2*)hi, (uint2*)lo);
  PT_LAUNCHED();
}

void split_f16_planes(const float* x, const float* add, int64_t n, int C, void* hi, void* lo, cudaStream_t s) {
  PT_CHECK(x && hi && lo && C >= 1, "split_f16: bad argument");
  if (n == 0) return;
  ProfScope prof(PROF_OTHER, s, 0.0, 8.0 * (double)n);
  split_f16_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(x, add, n, C, (__half*)hi, (__half*)lo);
  PT_LAUNCHED();
}

}  // namespace pttspp

extern "C" int pttspp_pack_conv_weight_split(const float* v, const float* g, int Cout, int Cin, int K, void* w_hi,
                                             void* w_lo, int interleave_halves, float* scale_inv) {
  PT_API_BEGIN
  PT_CHECK(v && w_hi && w_lo && scale_inv, "null argument");
  pttspp::pack_conv_weight_split(v, g, Cout, Cin, K, w_hi, w_lo, interleave_halves, scale_inv);
  PT_API_END
}

extern "C" int pttspp_split_f16(const float* x, const float* add, int64_t n, int C, void* hi, void* lo,
                                pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::split_f16_planes(x, add, n, C, hi, lo, (cudaStream_t)stream);
  PT_API_END
}

extern "C" int pttspp_conv1d_cl(const pttspp_conv1d_desc* d, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(d != nullptr, "null descriptor");
  pttspp::conv1d_cl(*d, (cudaStream_t)stream);
  PT_API_END
}

extern "C" int pttspp_conv1d_dual_cl(const pttspp_conv1d_desc* d1, const pttspp_conv1d_desc* d2, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(d1 && d2, "null descriptor");
  pttspp::conv1d_umma_dual_cl(*d1, *d2, (cudaStream_t)stream);
  PT_API_END
}

extern "C" int pttspp_pack_conv_weight(const float* v, const float* g, int Cout, int Cin, int K, float* packed,
                                       int w_ld, int interleave_halves, pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::pack_conv_weight(v, g, Cout, Cin, K, packed, w_ld, interleave_halves, (cudaStream_t)stream);
  PT_API_END
}

extern "C" int pttspp_pack_convtr_weight(const float* v, const float* g, int Cin, int Cout, int Kt, int stride,
                                         float* packed, int w_ld, pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::pack_convtr_weight(v, g, Cin, Cout, Kt, stride, packed, w_ld, (cudaStream_t)stream);
  PT_API_END
}
