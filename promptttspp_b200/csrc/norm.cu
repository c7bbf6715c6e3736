// LayerNorm over the channel dim of channels-last activations: one warp per row, the row lives
// in registers (two passes over registers: mean, then centred variance), float4 HBM access.
// Fused prologue: input length mask, scale, second addend, per-row table (sinusoidal PE);
// fused epilogue: affine + output length mask.  HBM-bound: 2*4*C bytes per row.
#include "common.h"

namespace pttspp {
namespace {

constexpr int LN_MAX_V4 = 8;  // C <= 8*32*4 = 1024

__global__ void __launch_bounds__(256) layernorm_kernel(const pttspp_layernorm_desc d) {
  const int warps_per_block = blockDim.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)d.B * d.T;
  if (row >= rows) return;
  const int b = (int)(row / d.T), t = (int)(row % d.T);
  const int64_t base = (int64_t)b * d.bs + (int64_t)t * d.ld;
  const int nv4 = d.C >> 2;
  const bool in_valid = !d.in_len || (long long)t < (long long)d.in_len[b];
  const bool out_valid = !d.out_len || (long long)t < (long long)d.out_len[b];

  float4 x[LN_MAX_V4];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_V4; ++j) {
    const int c4 = lane + 32 * j;
    x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < nv4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in_valid) v = *reinterpret_cast<const float4*>(d.in + base + c4 * 4);
      v.x *= d.in_scale; v.y *= d.in_scale; v.z *= d.in_scale; v.w *= d.in_scale;
      if (d.in2) {
        const float4 a = *reinterpret_cast<const float4*>(d.in2 + base + c4 * 4);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      if (d.row_add) {
        const float4 a = *reinterpret_cast<const float4*>(d.row_add + (int64_t)t * d.C + c4 * 4);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      x[j] = v;
      sum += (v.x + v.y) + (v.z + v.w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)d.C;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_V4; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < nv4) {
      const float a = x[j].x - mean, bq = x[j].y - mean, c = x[j].z - mean, e = x[j].w - mean;
      sq += (a * a + bq * bq) + (c * c + e * e);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.f / sqrtf(sq / (float)d.C + d.eps);
#pragma unroll
  for (int j = 0; j < LN_MAX_V4; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < nv4) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (out_valid) {
        const float4 g = *reinterpret_cast<const float4*>(d.gamma + c4 * 4);
        const float4 be = *reinterpret_cast<const float4*>(d.beta + c4 * 4);
        y.x = (x[j].x - mean) * rstd * g.x + be.x;
        y.y = (x[j].y - mean) * rstd * g.y + be.y;
        y.z = (x[j].z - mean) * rstd * g.z + be.z;
        y.w = (x[j].w - mean) * rstd * g.w + be.w;
      }
      *reinterpret_cast<float4*>(d.out + base + c4 * 4) = y;
    }
  }
}

}  // namespace

void layernorm_cl(const pttspp_layernorm_desc& d, cudaStream_t s) {
  PT_CHECK(d.in && d.out && d.gamma && d.beta, "layernorm: null pointer");
  PT_CHECK(d.C % 4 == 0 && d.C >= 4 && d.C <= LN_MAX_V4 * 128, "layernorm: C=%d unsupported", d.C);
  PT_CHECK(d.ld % 4 == 0 && d.bs % 4 == 0, "layernorm: strides must be multiples of 4");
  PT_CHECK(aligned16(d.in) && aligned16(d.out) && aligned16(d.gamma) && aligned16(d.beta) &&
               (!d.in2 || aligned16(d.in2)) && (!d.row_add || aligned16(d.row_add)),
           "layernorm: pointers must be 16-byte aligned");
  const int64_t rows = (int64_t)d.B * d.T;
  if (rows == 0) return;
  ProfScope prof(PROF_LAYERNORM, s, 0.0, 2.0 * 4.0 * (double)rows * d.C);
  const int wpb = 8;
  layernorm_kernel<<<(unsigned)ceil_div64(rows, wpb), wpb * 32, 0, s>>>(d);
  PT_LAUNCHED();
}

}  // namespace pttspp

extern "C" int pttspp_layernorm_cl(const pttspp_layernorm_desc* d, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(d != nullptr, "null descriptor");
  pttspp::layernorm_cl(*d, (cudaStream_t)stream);
  PT_API_END
}
