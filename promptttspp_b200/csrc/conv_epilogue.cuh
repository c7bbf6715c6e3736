// Fused conv epilogue shared by the CUDA-core (conv1d.cu) and tcgen05 (conv1d_umma.cu) kernels.
//
//   v   = act(acc_scale * acc + bias[c] + addend[row, c])          (gate: pairs of columns -> one output)
//   y   = (res_scale * res[row, oc] + alpha * mask * v + beta * out[row, oc]) / out_div
//   out[row, oc] = y                                               (skipped when out == NULL)
//   out_hi/out_lo[row, oc] = split_fp16(y + out_plane_add[oc])     (optional: operand planes for the next conv)
#pragma once
#include <cuda_fp16.h>

#include "common.h"

namespace pttspp {

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case PTTSPP_ACT_RELU: return fmaxf(v, 0.f);
    case PTTSPP_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case PTTSPP_ACT_SWISH: return v / (1.f + expf(-v));
    case PTTSPP_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = pt_f2h_sat(v);
  lo = pt_f2h_sat(v - __half2float(hi));
}

// Four consecutive pre-activation columns col..col+3 (col % 4 == 0) of output row `row` of batch `b`.
// `acc` are raw accumulators.  Columns >= Cout are ignored.
__device__ __forceinline__ void conv_epilogue4(const pttspp_conv1d_desc& d, int b, int row, float mask, int col,
                                               const float (&acc)[4]) {
  if (col >= d.Cout) return;
  const bool gate = (d.act == PTTSPP_ACT_GATE);
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = col + e;
    float t = acc[e] * d.acc_scale;
    if (c < d.Cout) {
      if (d.bias) t += d.bias[c];
      if (d.addend) t += d.addend[(int64_t)b * d.addend_bs + (int64_t)row * d.addend_ld + c];
    }
    v[e] = t;
  }
  float o[4];
  int nout, ocol, out_cols;
  if (gate) {
    o[0] = (1.f / (1.f + expf(-v[0]))) * tanhf(v[1]);
    o[1] = (1.f / (1.f + expf(-v[2]))) * tanhf(v[3]);
    o[2] = o[3] = 0.f;
    nout = 2;
    ocol = col >> 1;
    out_cols = d.Cout >> 1;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = act_apply(v[e], d.act);
    nout = 4;
    ocol = col;
    out_cols = d.Cout;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (e >= nout) break;
    const int oc = ocol + e;
    if (oc >= out_cols) break;
    float y = d.alpha * mask * o[e];
    if (d.res) y += d.res_scale * d.res[(int64_t)b * d.res_bs + (int64_t)row * d.res_ld + oc];
    if (d.out) {
      const int64_t oidx = (int64_t)b * d.out_bs + (int64_t)row * d.out_ld + oc;
      if (d.beta != 0.f) y += d.beta * d.out[oidx];
      if (d.out_div != 0.f) y = y / d.out_div;
      d.out[oidx] = y;
    } else if (d.out_div != 0.f) {
      y = y / d.out_div;
    }
    if (d.out_hi) {
      const float pv = d.out_plane_add ? y + d.out_plane_add[oc] : y;
      __half hi, lo;
      split_f16(pv, hi, lo);
      const int64_t pidx = (int64_t)b * d.out_plane_bs + (int64_t)row * d.out_plane_ld + oc;
      reinterpret_cast<__half*>(d.out_hi)[pidx] = hi;
      reinterpret_cast<__half*>(d.out_lo)[pidx] = lo;
    }
  }
}

// sigmoid(g) * tanh(f) on the MUFU path: ex2.approx (2^-22 rel) + one division; tanh(f) = (1 - e^-2f) / (1 + e^-2f)
// with f clamped to +-15 (tanh is +-1 to 1e-13 beyond).  ~1e-7 absolute: the same class as libm's expf/tanhf.
__device__ __forceinline__ float gate_fast(float g, float f) {
  const float fc = fminf(fmaxf(f, -15.f), 15.f);
  const float eg = __expf(-g);
  const float ef = __expf(-2.f * fc);
  return __fdividef(1.f - ef, (1.f + eg) * (1.f + ef));
}

// Vector variant for the tcgen05 kernel: one thread owns 16 consecutive pre-activation columns col..col+15
// (col % 16 == 0, col + 16 <= Cout) of one row and moves them with 16-byte accesses (every 32-byte sector it
// touches is fully used).  Preconditions (checked on the host, see conv_epilogue_vec_ok): all row strides are
// multiples of 4 floats / 8 halves and all base pointers are 16-byte aligned.
// Split in two phases so the kernel can issue the operand loads of the next chunk (or of the next tile, before its
// accumulator is even complete) while it works on the current one.
// One operand tile is prefetched per 16-column chunk (priority: addend, else residual, else previous output --
// the DiffNet launches each have exactly one of them); further operands are loaded in place.
struct EpiOps16 {
  float4 p[4];
};

// 1 conditioner addend, 2 fp32 residual, 3 previous output (beta), 4 residual from split-fp16 planes
__device__ __forceinline__ int conv_epilogue_prefetch_kind(const pttspp_conv1d_desc& d) {
  return d.addend ? 1 : (d.res ? 2 : (d.res_hi ? 4 : ((d.out && d.beta != 0.f) ? 3 : 0)));
}

__device__ __forceinline__ void conv_epilogue16_load(const pttspp_conv1d_desc& d, int b, int row, int col, EpiOps16& o) {
  const bool gate = (d.act == PTTSPP_ACT_GATE);
  const int ocol = gate ? (col >> 1) : col;
  const int kind = conv_epilogue_prefetch_kind(d);
  const float4* src = nullptr;
  if (kind == 1) src = reinterpret_cast<const float4*>(d.addend + (int64_t)b * d.addend_bs + (int64_t)row * d.addend_ld + col);
  else if (kind == 2) src = reinterpret_cast<const float4*>(d.res + (int64_t)b * d.res_bs + (int64_t)row * d.res_ld + ocol);
  else if (kind == 3) src = reinterpret_cast<const float4*>(d.out + (int64_t)b * d.out_bs + (int64_t)row * d.out_ld + ocol);
  const int nq = (kind == 1 || !gate) ? 4 : 2;
  if (src) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nq) o.p[i] = src[i];
  }
}

// arithmetic only: o[0..nout) = final output values of columns ocol.. (nout = 8 for the gate activation, else 16)
__device__ __forceinline__ void conv_epilogue16_math(const pttspp_conv1d_desc& d, int b, int row, float mask, int col,
                                                     float (&v)[16], const EpiOps16& ops, float (&o)[16]) {
  const bool gate = (d.act == PTTSPP_ACT_GATE);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] *= d.acc_scale;
  if (d.bias) {
    const float4* bp = reinterpret_cast<const float4*>(d.bias + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(bp + i);
      v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
  }
  const int pk = conv_epilogue_prefetch_kind(d);
  if (d.addend) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = ops.p[i];
      v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
  }
  const int nout = gate ? 8 : 16;
  const int ocol = gate ? (col >> 1) : col;
  if (gate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = gate_fast(v[2 * i], v[2 * i + 1]);
#pragma unroll
    for (int i = 8; i < 16; ++i) o[i] = 0.f;
  } else if (d.act == PTTSPP_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = v[i];
  } else if (d.act == PTTSPP_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = fmaxf(v[i], 0.f);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = act_apply(v[i], d.act);
  }
  const float am = d.alpha * mask;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < nout) o[i] *= am;
  if (d.res) {
    const float4* rp = reinterpret_cast<const float4*>(d.res + (int64_t)b * d.res_bs + (int64_t)row * d.res_ld + ocol);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * i < nout) {
        const float4 t = (pk == 2) ? ops.p[i] : rp[i];
        o[4 * i] += d.res_scale * t.x; o[4 * i + 1] += d.res_scale * t.y;
        o[4 * i + 2] += d.res_scale * t.z; o[4 * i + 3] += d.res_scale * t.w;
      }
  }
  if (d.out && d.beta != 0.f) {
    const float4* op = reinterpret_cast<const float4*>(d.out + (int64_t)b * d.out_bs + (int64_t)row * d.out_ld + ocol);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * i < nout) {
        const float4 t = (pk == 3) ? ops.p[i] : op[i];
        o[4 * i] += d.beta * t.x; o[4 * i + 1] += d.beta * t.y; o[4 * i + 2] += d.beta * t.z; o[4 * i + 3] += d.beta * t.w;
      }
  }
  if (d.out_div != 0.f) {
    const float inv = 1.f / d.out_div;  // <= 1 ulp from the IEEE division of the reference
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nout) o[i] = o[i] * inv;
  }
}

// direct (row-per-lane) stores of the values computed by conv_epilogue16_math
__device__ __forceinline__ void conv_epilogue16_store(const pttspp_conv1d_desc& d, int b, int row, int col, float (&o)[16]) {
  const bool gate = (d.act == PTTSPP_ACT_GATE);
  const int nout = gate ? 8 : 16;
  const int ocol = gate ? (col >> 1) : col;
  if (d.out) {
    float4* op = reinterpret_cast<float4*>(d.out + (int64_t)b * d.out_bs + (int64_t)row * d.out_ld + ocol);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * i < nout) op[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
  if (d.out_hi) {
    if (d.out_plane_add) {
      const float4* pp = reinterpret_cast<const float4*>(d.out_plane_add + ocol);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (4 * i < nout) {
          const float4 t = __ldg(pp + i);
          o[4 * i] += t.x; o[4 * i + 1] += t.y; o[4 * i + 2] += t.z; o[4 * i + 3] += t.w;
        }
    }
    const int64_t pidx = (int64_t)b * d.out_plane_bs + (int64_t)row * d.out_plane_ld + ocol;
    uint4* hp = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(d.out_hi) + pidx);
    uint4* lp = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(d.out_lo) + pidx);
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (8 * i < nout) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __half h0, l0, h1, l1;
          split_f16(o[8 * i + 2 * j], h0, l0);
          split_f16(o[8 * i + 2 * j + 1], h1, l1);
          hw[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          lw[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        hp[i] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        lp[i] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
  }
}

__device__ __forceinline__ void conv_epilogue16_finish(const pttspp_conv1d_desc& d, int b, int row, float mask, int col,
                                                       float (&v)[16], const EpiOps16& ops) {
  float o[16];
  conv_epilogue16_math(d, b, row, mask, col, v, ops, o);
  conv_epilogue16_store(d, b, row, col, o);
}

// host-side precondition of the conv_epilogue16_* vector path
static inline bool conv_epilogue_vec_ok(const pttspp_conv1d_desc& d) {
  auto ok4 = [](const void* p, int64_t bs, int ld) { return !p || (aligned16(p) && bs % 4 == 0 && ld % 4 == 0); };
  if (d.Cout % 16 != 0) return false;
  if (d.bias && !aligned16(d.bias)) return false;
  if (!ok4(d.addend, d.addend_bs, d.addend_ld) || !ok4(d.res, d.res_bs, d.res_ld) || !ok4(d.out, d.out_bs, d.out_ld))
    return false;
  if (d.out_hi && !(aligned16(d.out_hi) && aligned16(d.out_lo) && d.out_plane_bs % 8 == 0 && d.out_plane_ld % 8 == 0))
    return false;
  if (d.out_plane_add && !aligned16(d.out_plane_add)) return false;
  return true;
}

}  // namespace pttspp
