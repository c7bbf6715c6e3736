// Relative-position multi-head self-attention of the Conformer text encoder (Transformer-XL
// style), both ESPnet variants, CUDA-core fp32 path.
//
//   ac[i,j]  = (q_i + u) . k_j                       bd[i,j'] = (q_i + v) . p_j'
//   new:     bd'[i,j] = bd[i, T-1+j-i]                               (Tp = 2T-1)
//   legacy:  bd'[i,j] = bd[i, T-1-i+j]   if j <= i                   (Tp = T)
//                      = 0               if j == i+1
//                      = bd[i+1, j-i-2]  if j >= i+2   (the wrapped upper triangle the legacy
//                                                       rel_shift leaves behind -- it feeds the softmax)
//   scores = (ac + bd') / sqrt(d_k);  keys j >= len and queries i >= len are masked
//   (softmax over valid keys; a fully masked query row yields 0, attention.py:77-84).
//
// Two launches: (1) bd for all (b,h) into scratch (a small tiled GEMM), (2) one CTA per
// (b, h, 16-query tile): scores -> shared memory, warp softmax, P.V with V streamed through
// shared memory.  The problem is 32 heads of 256x256x128 -- latency bound; fusion is the lever.
#include "common.h"

namespace pttspp {
namespace {

constexpr int TI = 16;   // queries per CTA
constexpr int TJ = 32;   // keys per streamed tile
constexpr int MAX_E = 8; // d_k <= 256

// bd[b][h][i][j'] ; grid (ceil(Tp/32), ceil(T/32), B*H), 256 threads
__global__ void __launch_bounds__(256) relpos_bd_kernel(const float* __restrict__ q, const float* __restrict__ p,
                                                        const float* __restrict__ bias_v, int T, int Tp, int H,
                                                        int dk, int ld, float* __restrict__ bd) {
  extern __shared__ float sm[];
  const int ldk = dk + 1;
  float* Qs = sm;             // [32][dk+1]
  float* Ps = sm + 32 * ldk;  // [32][dk+1]
  const int bh = blockIdx.z, b = bh / H, h = bh % H;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int idx = threadIdx.x; idx < 32 * dk; idx += blockDim.x) {
    const int r = idx / dk, dd = idx % dk;
    const int i = i0 + r, j = j0 + r;
    Qs[r * ldk + dd] = (i < T) ? q[((int64_t)b * T + i) * ld + h * dk + dd] + bias_v[h * dk + dd] : 0.f;
    Ps[r * ldk + dd] = (j < Tp) ? p[(int64_t)j * (H * dk) + h * dk + dd] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int dd = 0; dd < dk; ++dd) {
    const float pv = Ps[tx * ldk + dd];
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = fmaf(Qs[(ty + 8 * e) * ldk + dd], pv, acc[e]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = i0 + ty + 8 * e, j = j0 + tx;
    if (i < T && j < Tp) bd[((int64_t)bh * T + i) * Tp + j] = acc[e];
  }
}

// grid (ceil(T/TI), H, B), 256 threads (8 warps x 2 query rows)
__global__ void __launch_bounds__(256) relpos_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                          const float* __restrict__ v,
                                                          const float* __restrict__ bias_u,
                                                          const float* __restrict__ bd,
                                                          const int64_t* __restrict__ lens, int T, int Tp, int H,
                                                          int dk, int ld, int legacy, float scale,
                                                          float* __restrict__ out) {
  extern __shared__ float sm[];
  const int ldk = dk + 1;
  float* S = sm;                 // [TI][T]
  float* Qs = S + TI * T;        // [TI][dk]
  float* KV = Qs + TI * dk;      // [TJ][dk+1]
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TI;
  const int bh = b * H + h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long l64 = lens ? lens[b] : (long long)T;
  const int len = (int)(l64 < (long long)T ? l64 : (long long)T);
  const int HD = H * dk;

  for (int idx = threadIdx.x; idx < TI * dk; idx += blockDim.x) {
    const int r = idx / dk, dd = idx % dk;
    const int i = i0 + r;
    Qs[idx] = (i < T) ? q[((int64_t)b * T + i) * ld + h * dk + dd] + bias_u[h * dk + dd] : 0.f;
  }
  // ---- scores ----
  for (int j0 = 0; j0 < len; j0 += TJ) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TJ * dk; idx += blockDim.x) {
      const int r = idx / dk, dd = idx % dk;
      const int j = j0 + r;
      KV[r * ldk + dd] = (j < len) ? k[((int64_t)b * T + j) * ld + h * dk + dd] : 0.f;
    }
    __syncthreads();
    const int j = j0 + lane;
    float a0 = 0.f, a1 = 0.f;
    const float* q0 = Qs + (2 * warp) * dk;
    const float* q1 = q0 + dk;
    const float* kr = KV + lane * ldk;
    for (int dd = 0; dd < dk; ++dd) {
      const float kv = kr[dd];
      a0 = fmaf(q0[dd], kv, a0);
      a1 = fmaf(q1[dd], kv, a1);
    }
    if (j < len) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 2 * warp + e, i = i0 + r;
        if (i < T) {
          float pos;
          const float* bdb = bd + (int64_t)bh * T * Tp;
          if (!legacy) {
            pos = bdb[(int64_t)i * Tp + (T - 1 + j - i)];
          } else if (j <= i) {
            pos = bdb[(int64_t)i * Tp + (T - 1 - i + j)];
          } else if (j == i + 1) {
            pos = 0.f;
          } else {
            pos = bdb[(int64_t)(i + 1) * Tp + (j - i - 2)];
          }
          S[r * T + j] = ((e == 0 ? a0 : a1) + pos) * scale;
        }
      }
    }
  }
  __syncthreads();
  // ---- softmax over valid keys (warp per 2 rows) ----
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int r = 2 * warp + e, i = i0 + r;
    if (i < len) {
      float mx = -INFINITY;
      for (int j = lane; j < len; j += 32) mx = fmaxf(mx, S[r * T + j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int j = lane; j < len; j += 32) {
        const float ev = expf(S[r * T + j] - mx);
        S[r * T + j] = ev;
        sum += ev;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      for (int j = lane; j < len; j += 32) S[r * T + j] *= inv;
    }
  }
  // ---- O = P . V ----
  float acc[2][MAX_E];
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int c = 0; c < MAX_E; ++c) acc[e][c] = 0.f;
  const int ne = dk >> 5;
  for (int j0 = 0; j0 < len; j0 += TJ) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TJ * dk; idx += blockDim.x) {
      const int r = idx / dk, dd = idx % dk;
      const int j = j0 + r;
      KV[r * ldk + dd] = (j < len) ? v[((int64_t)b * T + j) * ld + h * dk + dd] : 0.f;
    }
    __syncthreads();
    const int jn = (len - j0 < TJ) ? (len - j0) : TJ;
    for (int jj = 0; jj < jn; ++jj) {
      const float p0 = S[(2 * warp) * T + j0 + jj];
      const float p1 = S[(2 * warp + 1) * T + j0 + jj];
#pragma unroll
      for (int c = 0; c < MAX_E; ++c) {
        if (c < ne) {
          const float vv = KV[jj * ldk + lane + 32 * c];
          acc[0][c] = fmaf(p0, vv, acc[0][c]);
          acc[1][c] = fmaf(p1, vv, acc[1][c]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = i0 + 2 * warp + e;
    if (i < T) {
      const bool valid = i < len;
#pragma unroll
      for (int c = 0; c < MAX_E; ++c)
        if (c < ne) out[((int64_t)b * T + i) * HD + h * dk + lane + 32 * c] = valid ? acc[e][c] : 0.f;
    }
  }
}

}  // namespace

bool relpos_attention_umma(const float* q, const float* k, const float* v, const float* p, const float* bias_u,
                           const float* bias_v, const int64_t* lens, int B, int T, int H, int dk, int legacy, float* out,
                           int ld_qkv, cudaStream_t s);  // attention_umma.cu

void relpos_attention(const float* q, const float* k, const float* v, const float* p, const float* bias_u,
                      const float* bias_v, const int64_t* lens, int B, int T, int H, int dk, int legacy,
                      float* scratch, float* out, int ld_qkv, cudaStream_t s) {
  PT_CHECK(q && k && v && p && bias_u && bias_v && scratch && out, "relpos_attention: null pointer");
  PT_CHECK(dk % 32 == 0 && dk <= 32 * MAX_E, "relpos_attention: d_k=%d unsupported", dk);
  PT_CHECK(B * H <= 65535 && B <= 65535, "relpos_attention: batch too large");
  if (B == 0 || T == 0) return;
  const int Tp = legacy ? T : 2 * T - 1;
  ProfScope prof(PROF_ATTENTION, s, 2.0 * B * H * (double)T * dk * (2.0 * T + Tp), 4.0 * 4.0 * B * (double)T * H * dk);
  // fused tcgen05 kernel (d_k = 128, T <= 256): no scratch matrix, one launch
  if (relpos_attention_umma(q, k, v, p, bias_u, bias_v, lens, B, T, H, dk, legacy, out, ld_qkv, s)) return;
  {
    const size_t smem = (size_t)2 * 32 * (dk + 1) * sizeof(float);
    if (smem > 48 * 1024)
      PT_CUDA(cudaFuncSetAttribute(relpos_bd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(Tp, 32), ceil_div(T, 32), B * H);
    relpos_bd_kernel<<<grid, 256, smem, s>>>(q, p, bias_v, T, Tp, H, dk, ld_qkv, scratch);
    PT_LAUNCHED();
  }
  {
    const size_t smem = ((size_t)TI * T + (size_t)TI * dk + (size_t)TJ * (dk + 1)) * sizeof(float);
    PT_CHECK(smem <= 220 * 1024, "relpos_attention: T=%d too long for the shared-memory score tile", T);
    if (smem > 48 * 1024)
      PT_CUDA(cudaFuncSetAttribute(relpos_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(T, TI), H, B);
    relpos_attn_kernel<<<grid, 256, smem, s>>>(q, k, v, bias_u, scratch, lens, T, Tp, H, dk, ld_qkv, legacy,
                                               1.f / sqrtf((float)dk), out);
    PT_LAUNCHED();
  }
}

}  // namespace pttspp

extern "C" int pttspp_relpos_attention(const float* q, const float* k, const float* v, const float* p,
                                       const float* bias_u, const float* bias_v, const int64_t* lens, int B, int T,
                                       int H, int dk, int legacy, float* scratch, float* out,
                                       pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::relpos_attention(q, k, v, p, bias_u, bias_v, lens, B, T, H, dk, legacy, scratch, out, H * dk,
                           (cudaStream_t)stream);
  PT_API_END
}
