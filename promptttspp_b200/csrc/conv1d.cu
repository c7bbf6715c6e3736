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

namespace pttspp {

namespace {

constexpr int BK = 16;
constexpr int LDA = BK + 4;

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case PTTSPP_ACT_RELU: return fmaxf(v, 0.f);
    case PTTSPP_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case PTTSPP_ACT_SWISH: return v / (1.f + expf(-v));
    case PTTSPP_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

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

  // ---- epilogue ----
  const bool gate = (d.act == PTTSPP_ACT_GATE);
  const int out_cols = gate ? d.Cout / 2 : d.Cout;
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
      const int col = n0 + g * GS + tx * 4;
      if (col >= d.Cout) continue;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = col + e;
        float t = acc[i][g * 4 + e] * d.acc_scale;
        if (c < d.Cout) {
          if (d.bias) t += d.bias[c];
          if (d.addend) t += d.addend[(int64_t)b * d.addend_bs + (int64_t)row * d.addend_ld + c];
        }
        v[e] = t;
      }
      float o[4];
      int nout, ocol;
      if (gate) {
        o[0] = (1.f / (1.f + expf(-v[0]))) * tanhf(v[1]);
        o[1] = (1.f / (1.f + expf(-v[2]))) * tanhf(v[3]);
        o[2] = o[3] = 0.f;
        nout = 2;
        ocol = col >> 1;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = act_apply(v[e], d.act);
        nout = 4;
        ocol = col;
      }
      for (int e = 0; e < nout; ++e) {
        const int oc = ocol + e;
        if (oc >= out_cols) break;
        const int64_t oidx = (int64_t)b * d.out_bs + (int64_t)row * d.out_ld + oc;
        float y = d.alpha * mask * o[e];
        if (d.res) y += d.res_scale * d.res[(int64_t)b * d.res_bs + (int64_t)row * d.res_ld + oc];
        if (d.beta != 0.f) y += d.beta * d.out[oidx];
        if (d.out_div != 0.f) y = y / d.out_div;
        d.out[oidx] = y;
      }
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
  PT_CHECK(d.in && d.w && d.out, "conv1d: null pointer");
  PT_CHECK(d.Cin > 0 && d.Cin % BK == 0, "conv1d: Cin=%d must be a positive multiple of %d", d.Cin, BK);
  PT_CHECK(d.in_ld % 4 == 0 && aligned16(d.in), "conv1d: input must be 16-byte aligned, ld %% 4 == 0");
  PT_CHECK(d.w_ld % 4 == 0 && d.w_ld >= d.Cout && aligned16(d.w), "conv1d: packed weight ld=%d invalid", d.w_ld);
  PT_CHECK(!d.in_add || aligned16(d.in_add), "conv1d: in_add must be 16-byte aligned");
  PT_CHECK(d.K >= 1 && d.dil >= 1 && d.in_stride >= 1 && d.out_mul >= 1, "conv1d: bad geometry");
  PT_CHECK(d.act != PTTSPP_ACT_GATE || d.Cout % 2 == 0, "conv1d: gate activation needs even Cout");
  PT_CHECK(d.B >= 1 && d.B <= 65535, "conv1d: batch %d out of range", d.B);
  if (d.M <= 0 || d.Cout <= 0) return;
  const double flops = 2.0 * d.B * (double)d.M * d.Cout * (double)d.Cin * d.K;
  const bool umma = (d.impl == 2) || (d.impl == 0 && conv1d_umma_supported(d));
  ProfScope prof(umma ? PROF_CONV_UMMA : PROF_CONV_SIMT, s, flops, 0.0);
  if (d.impl == 2) {
    PT_CHECK(conv1d_umma_supported(d), "conv1d: tcgen05 path requested for an unsupported shape");
    conv1d_umma_cl(d, s);
    return;
  }
  if (d.impl == 0 && conv1d_umma_supported(d)) {
    conv1d_umma_cl(d, s);
    return;
  }
  if (d.Cout > 64)
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

}  // namespace pttspp

extern "C" int pttspp_conv1d_cl(const pttspp_conv1d_desc* d, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(d != nullptr, "null descriptor");
  pttspp::conv1d_cl(*d, (cudaStream_t)stream);
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
