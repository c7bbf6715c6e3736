// Packed fp32x2 arithmetic of the anti-aliased Snake activation (layers/activations.py:22-138), shared by the
// stand-alone channel-pair kernel (aa_snake.cu) and the activation-producer warps of the fused AA -> conv kernel
// (conv1d_umma.cu).  ncu: FFMA2 occupies the FP32 pipe for two cycles per warp, so the packed forms save issue
// slots, not pipe time.
#pragma once
#include <cuda_fp16.h>

namespace pttspp {
namespace {

typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// s = u + sin^2(alpha * u) / (alpha + 1e-9), both channels (same operations as the strip kernel, packed where possible)
__device__ __forceinline__ f32x2 snake2(f32x2 u, f32x2 a2, f32x2 inv_alpha) {
  float x0, x1;
  upk2(mul2(u, a2), x0, x1);
  const f32x2 sn = pk2(__sinf(x0), __sinf(x1));
  return fma2(mul2(inv_alpha, sn), sn, u);
}

}  // namespace
}  // namespace pttspp
