// tcgen05 (UMMA) implicit-GEMM Conv1d on split-fp16 operand planes, fed by TMA.
//
// An fp32 activation / weight v travels as two fp16 planes (hi = fp16(v), lo = fp16(v - hi)).  Per K-slab the
// tensor cores run three kind::f16 MMAs with fp32 accumulation in TMEM:
//        D += A_hi * B_hi;   D += A_hi * B_lo;   D += A_lo * B_hi        (dropped term lo*lo ~ 2^-22)
// which keeps the contraction in the fp32 error class (the parity bar is 1e-3 on mel after 100 x 20 layers; single
// bf16/tf32 MMAs do not meet it) at 3 tensor-core passes -- 1.5x the cost of one tf32 pass.
//
// Implicit GEMM: time is MMA-M (128 rows per CTA), output channels are MMA-N (BN = 128 per CTA), K = taps x Cin.
// For every (64-channel slab, tap) the producer issues four TMA tile loads into one pipeline stage:
//   A_hi/A_lo : box {64 ch, 128 rows} of the [B][T][Cin] planes at row  m0 + tap*dil - pad   (rows outside [0,T)
//               are zero-filled by TMA = the conv's zero padding, with no cross-batch bleed: 3-D tensor map)
//   B_hi/B_lo : box {64 ch, BN rows}  of the [taps*Cout][Cin] weight planes at row tap*Cout + n0
// all in the canonical K-major SWIZZLE_128B layout that the UMMA shared-memory descriptors expect.
//
// Warp roles (320 threads): warps 0-7 epilogue (TMEM -> registers -> fused epilogue -> HBM), warp 8 TMA producer
// (one elected lane), warp 9 TMEM allocator + MMA issuer (one elected lane).  Full/empty mbarriers form a
// 3-stage ring; tcgen05.commit releases a stage when its MMAs retire and finally signals the epilogue.
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>
#include <unordered_map>

#include "aa_math.cuh"
#include "common.h"
#include "conv_epilogue.cuh"
#include "umma_ptx.cuh"

namespace pttspp {
namespace {


// ---- coalescing epilogue -------------------------------------------------------------------------------------------
// The accumulator leaves TMEM row-per-lane (lane = tile row), crosses a 2 KB per-warp staging tile (XOR-swizzled:
// conflict-free both ways) and is finished in a row-coalesced layout: lane = (row lane/4 of an 8-row group, 16-byte
// column quad lane%4), so every global access of a warp instruction covers 8 rows x 64 contiguous bytes (full
// sectors) and nothing waits on the async proxy.  Operand tiles (conditioner / residual / previous output) are
// requested in the same layout before the accumulator is complete.
// `de` must be one of the kernel's by-value descriptor parameters, selected by a BRANCH at the call site (not by
// `cond ? d2 : d`): then every field is an immediate constant-bank operand; a run-time selected reference makes
// each access an indexed LDC with a long-scoreboard wait (that was most of the old epilogue's time).
// Precondition (host: epilogue_co_ok): the vector-path alignment rules, and no residual / beta with the gate.
template <int BN, int NACC, int EW>
__device__ __forceinline__ void umma_tile_epilogue_co(const pttspp_conv1d_desc& de, int n0, int mt, int b, int u, uint32_t tparity,
                                                      int warp, int lane, uint32_t tmem_base, uint32_t tfull,
                                                      int n_main, float* stage, int dbg) {
  const int q = warp & 3, cgrp = warp >> 2;
  const int m0 = de.m_begin + mt * UM_BM;
  constexpr int CW = BN / (EW / 4);
  constexpr int NCH = CW / 16;
  const int cbeg = cgrp * CW;
  const bool gate = (de.act == PTTSPP_ACT_GATE);
  const int rsub = lane >> 2, cq = lane & 3;
  const int kind = conv_epilogue_prefetch_kind(de);
  const long long omask_len = de.out_len ? (long long)de.out_len[b] : (1ll << 62);
  // rows of this lane in the coalesced layout
  int rows[4];
  bool rok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = m0 + q * 32 + 8 * j + rsub;
    rows[j] = m * de.out_mul + de.out_off;
    rok[j] = (m < de.m_begin + de.M) && rows[j] >= 0 && rows[j] < de.T_out;
  }
  float4 pre[NCH][4];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = n0 + cbeg + 16 * c + 4 * cq;  // first pre-activation column of this lane
    const int ocol = gate ? (col >> 1) : col;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pre[c][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rok[j] && col < de.Cout && !(dbg & 4)) {
        if (kind == 1)
          pre[c][j] = *reinterpret_cast<const float4*>(de.addend + (int64_t)b * de.addend_bs + (int64_t)rows[j] * de.addend_ld + col);
        else if (kind == 2)
          pre[c][j] = *reinterpret_cast<const float4*>(de.res + (int64_t)b * de.res_bs + (int64_t)rows[j] * de.res_ld + ocol);
        else if (kind == 3)
          pre[c][j] = *reinterpret_cast<const float4*>(de.out + (int64_t)b * de.out_bs + (int64_t)rows[j] * de.out_ld + ocol);
      }
    }
  }
  // per-column operands of this lane (same columns for all its rows), requested before the accumulator is complete
  float4 bias_c[NCH], padd_c[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = n0 + cbeg + 16 * c + 4 * cq;
    const int ocol = gate ? (col >> 1) : col;
    bias_c[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    padd_c[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < de.Cout) {
      if (de.bias) bias_c[c] = __ldg(reinterpret_cast<const float4*>(de.bias + col));
      if (de.out_hi && de.out_plane_add) {
        if (gate) {
          const float2 t2 = __ldg(reinterpret_cast<const float2*>(de.out_plane_add + ocol));
          padd_c[c].x = t2.x; padd_c[c].y = t2.y;
        } else {
          padd_c[c] = __ldg(reinterpret_cast<const float4*>(de.out_plane_add + ocol));
        }
      }
    }
  }
  mbar_wait_warp(tfull, tparity);
  tc_fence_after();
  if (dbg & 64) return;  // experiment: mainloop only
  const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * NACC * BN);
  float4* st4 = reinterpret_cast<float4*>(stage);
  const float inv_div = (de.out_div != 0.f) ? 1.f / de.out_div : 1.f;  // <= 1 ulp from the reference's IEEE division
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = cbeg + 16 * c;
    if (n0 + col0 >= de.Cout) continue;  // warp-uniform
    float v[16];
    {
      // all accumulators of this chunk are requested back to back, one wait
      uint32_t acc[NACC][16];
      if (dbg & 32) {  // experiment: no TMEM reads
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[a][e] = 0u;
      } else {
        tmem_ld16_nowait(tbase + (uint32_t)((NACC - 1) * BN + col0), acc[NACC - 1]);
#pragma unroll
        for (int a = 0; a < NACC - 1; ++a)
          if (a < n_main) tmem_ld16_nowait(tbase + (uint32_t)(a * BN + col0), acc[a]);
        tmem_wait_ld();
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(acc[NACC - 1][e]);
#pragma unroll
      for (int a = 0; a < NACC - 1; ++a)
        if (a < n_main) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(acc[a][e]);
        }
    }
    __syncwarp();  // the previous chunk's reads of the staging tile are done
#pragma unroll
    for (int k = 0; k < 4; ++k)
      st4[lane * 4 + (k ^ ((lane >> 1) & 3))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    __syncwarp();
    const int col = n0 + col0 + 4 * cq;
    const int ocol = gate ? (col >> 1) : col;
    const float4 bias4 = bias_c[c], padd = padd_c[c];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 8 * j + rsub;
      const float4 a = st4[r * 4 + (cq ^ ((r >> 1) & 3))];
      if (!rok[j]) continue;
      const float mask = ((long long)rows[j] < omask_len) ? 1.f : 0.f;
      float x0 = a.x * de.acc_scale + bias4.x, x1 = a.y * de.acc_scale + bias4.y;
      float x2 = a.z * de.acc_scale + bias4.z, x3 = a.w * de.acc_scale + bias4.w;
      if (kind == 1) { x0 += pre[c][j].x; x1 += pre[c][j].y; x2 += pre[c][j].z; x3 += pre[c][j].w; }
      const float am = de.alpha * mask;
      if (gate) {
        float o0 = gate_fast(x0, x1) * am, o1 = gate_fast(x2, x3) * am;
        o0 *= inv_div; o1 *= inv_div;
        if (de.out && !(dbg & 8))
          *reinterpret_cast<float2*>(de.out + (int64_t)b * de.out_bs + (int64_t)rows[j] * de.out_ld + ocol) = make_float2(o0, o1);
        if (de.out_hi && !(dbg & 8)) {
          uint32_t hw, lw;
          split2_f16(o0 + padd.x, o1 + padd.y, hw, lw);
          const int64_t pidx = (int64_t)b * de.out_plane_bs + (int64_t)rows[j] * de.out_plane_ld + ocol;
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(de.out_hi) + pidx) = hw;
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(de.out_lo) + pidx) = lw;
        }
      } else {
        float o0, o1, o2, o3;
        if (de.act == PTTSPP_ACT_NONE) { o0 = x0; o1 = x1; o2 = x2; o3 = x3; }
        else if (de.act == PTTSPP_ACT_RELU) { o0 = fmaxf(x0, 0.f); o1 = fmaxf(x1, 0.f); o2 = fmaxf(x2, 0.f); o3 = fmaxf(x3, 0.f); }
        else { o0 = act_apply(x0, de.act); o1 = act_apply(x1, de.act); o2 = act_apply(x2, de.act); o3 = act_apply(x3, de.act); }
        if (am != 1.f) { o0 *= am; o1 *= am; o2 *= am; o3 *= am; }
        if (de.res) {
          float4 rv = pre[c][j];
          if (kind != 2) rv = *reinterpret_cast<const float4*>(de.res + (int64_t)b * de.res_bs + (int64_t)rows[j] * de.res_ld + ocol);
          if (de.res_scale == 1.f) { o0 += rv.x; o1 += rv.y; o2 += rv.z; o3 += rv.w; }
          else { o0 += de.res_scale * rv.x; o1 += de.res_scale * rv.y; o2 += de.res_scale * rv.z; o3 += de.res_scale * rv.w; }
        }
        if (de.out && de.beta != 0.f) {
          float4 ov = pre[c][j];
          if (kind != 3) ov = *reinterpret_cast<const float4*>(de.out + (int64_t)b * de.out_bs + (int64_t)rows[j] * de.out_ld + ocol);
          if (de.beta == 1.f) { o0 += ov.x; o1 += ov.y; o2 += ov.z; o3 += ov.w; }
          else { o0 += de.beta * ov.x; o1 += de.beta * ov.y; o2 += de.beta * ov.z; o3 += de.beta * ov.w; }
        }
        if (de.out_div != 0.f) { o0 *= inv_div; o1 *= inv_div; o2 *= inv_div; o3 *= inv_div; }
        if (de.out && !(dbg & 8))
          *reinterpret_cast<float4*>(de.out + (int64_t)b * de.out_bs + (int64_t)rows[j] * de.out_ld + ocol) = make_float4(o0, o1, o2, o3);
        if (de.out_hi && !(dbg & 8)) {
          uint32_t h01, l01, h23, l23;
          split2_f16(o0 + padd.x, o1 + padd.y, h01, l01);
          split2_f16(o2 + padd.z, o3 + padd.w, h23, l23);
          const int64_t pidx = (int64_t)b * de.out_plane_bs + (int64_t)rows[j] * de.out_plane_ld + ocol;
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(de.out_hi) + pidx) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(de.out_lo) + pidx) = make_uint2(l01, l23);
        }
      }
    }
  }
}

// ---- row-per-lane epilogue with 256-bit global accesses ------------------------------------------------------------
// The accumulator is finished in its TMEM layout (lane = tile row).  Every global access is one full 32-byte sector
// per lane (LDG/STG.256: 8 floats or 16 halves), so there is no transposition through shared memory: no staging
// tile (the pair kernel's shared memory goes to more weight stages) and no STS/LDS traffic competing with the
// tensor core's operand reads, which already take most of the shared-memory bandwidth.
// Precondition (host: epilogue_rl_ok): 32-byte aligned bases and row strides, Cout % 16 == 0; gate: no residual/beta.

// operand tile of the row-per-lane epilogue: the 16 floats of (row, 16-column chunk `col`) of the conditioner
// (kind 1), the residual (kind 2) or the previous output (kind 3); zeros when there is none or the row is outside
__device__ __forceinline__ void epi_rl_load_operand(const pttspp_conv1d_desc& de, int b, int row, bool ok, int col,
                                                    f8 (&dst)[2]) {
  const int kind = conv_epilogue_prefetch_kind(de);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[k].v[e] = 0.f;
  }
  if (ok && kind == 4 && col < de.Cout) {
    // 16 halves of each plane = one 256-bit load each; kept as raw bits, decoded by the consumer
    const int64_t pidx = (int64_t)b * de.res_plane_bs + (int64_t)row * de.res_plane_ld + col;
    dst[0] = ldg256(reinterpret_cast<const float*>(reinterpret_cast<const __half*>(de.res_hi) + pidx));
    dst[1] = ldg256(reinterpret_cast<const float*>(reinterpret_cast<const __half*>(de.res_lo) + pidx));
  } else if (ok && kind != 0 && col < de.Cout) {
    const float* src = (kind == 1) ? de.addend + (int64_t)b * de.addend_bs + (int64_t)row * de.addend_ld
                     : (kind == 2) ? de.res + (int64_t)b * de.res_bs + (int64_t)row * de.res_ld
                                   : de.out + (int64_t)b * de.out_bs + (int64_t)row * de.out_ld;
    dst[0] = ldg256(src + col);  // kinds 2/3 are excluded for the gate (host check): output column == col
    dst[1] = ldg256(src + col + 8);
  }
}
// L2 prefetch of the NEXT tile's operand rows by the epilogue warps themselves (one 128-byte line per lane and
// 32-column slice: no registers, no TMA queue), so that DRAM latency is paid one tile ahead
template <int BN, int EW>
__device__ __forceinline__ void epi_rl_prefetch_tile(const pttspp_conv1d_desc& de, int n0, int mt, int b, int warp, int lane) {
  const int kind = conv_epilogue_prefetch_kind(de);
  if (kind == 0) return;
  constexpr int CW = BN / (EW / 4);
  const int q = warp & 3, cbeg = (warp >> 2) * CW;
  const int m = de.m_begin + mt * UM_BM + q * 32 + lane;
  const int row = m * de.out_mul + de.out_off;
  if (!((m < de.m_begin + de.M) && row >= 0 && row < de.T_out)) return;
  if (kind == 4) {  // residual planes: 32 columns = 64 bytes of each plane
    const int64_t pidx = (int64_t)b * de.res_plane_bs + (int64_t)row * de.res_plane_ld + n0 + cbeg;
#pragma unroll
    for (int c = 0; c < CW; c += 32)
      if (n0 + cbeg + c < de.Cout) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __half*>(de.res_hi) + pidx + c));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __half*>(de.res_lo) + pidx + c));
      }
    return;
  }
  const float* src = (kind == 1) ? de.addend + (int64_t)b * de.addend_bs + (int64_t)row * de.addend_ld
                   : (kind == 2) ? de.res + (int64_t)b * de.res_bs + (int64_t)row * de.res_ld
                                 : de.out + (int64_t)b * de.out_bs + (int64_t)row * de.out_ld;
#pragma unroll
  for (int c = 0; c < CW; c += 32)
    if (n0 + cbeg + c < de.Cout) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + n0 + cbeg + c));
}

// (A rolling prefetch of the NEXT tile's operand chunks into the registers of consumed chunks was tried and measured
// slower: the longer live ranges spill at the 96-register budget of the 18-warp CTA.)
template <int BN, int NACC, int EW>
__device__ __forceinline__ void umma_tile_epilogue_rl(const pttspp_conv1d_desc& de, int n0, int mt, int b, int u,
                                                      uint32_t tparity, int warp, int lane, uint32_t tmem_base,
                                                      uint32_t tfull, int n_main, int dbg) {
  const int q = warp & 3, cgrp = warp >> 2;
  constexpr int CW = BN / (EW / 4);
  constexpr int NCH = CW / 16;
  const int cbeg = cgrp * CW;
  const bool gate = (de.act == PTTSPP_ACT_GATE);
  const int kind = conv_epilogue_prefetch_kind(de);
  const int m = de.m_begin + mt * UM_BM + q * 32 + lane;
  const int row = m * de.out_mul + de.out_off;
  const bool ok = (m < de.m_begin + de.M) && row >= 0 && row < de.T_out;
  float am = de.alpha;
  if (de.out_len && ok) am = ((long long)row < (long long)de.out_len[b]) ? de.alpha : 0.f;
  const float inv_div = (de.out_div != 0.f) ? 1.f / de.out_div : 1.f;  // <= 1 ulp from the reference's IEEE division
  // operand tile (conditioner / residual / previous output): 16 floats per chunk, requested before the accumulator
  // is complete
  f8 pre[NCH][2];
#pragma unroll
  for (int c = 0; c < NCH; ++c) epi_rl_load_operand(de, b, row, ok && !(dbg & 4), n0 + cbeg + 16 * c, pre[c]);
  mbar_wait_warp(tfull, tparity);
  tc_fence_after();
  if (dbg & 64) return;
  const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * NACC * BN);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = cbeg + 16 * c;
    if (n0 + col0 >= de.Cout) continue;  // warp-uniform
    float v[16];
    {
      uint32_t acc[NACC][16];
      tmem_ld16_nowait(tbase + (uint32_t)((NACC - 1) * BN + col0), acc[NACC - 1]);
#pragma unroll
      for (int a = 0; a < NACC - 1; ++a)
        if (a < n_main) tmem_ld16_nowait(tbase + (uint32_t)(a * BN + col0), acc[a]);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(acc[NACC - 1][e]);
#pragma unroll
      for (int a = 0; a < NACC - 1; ++a)
        if (a < n_main) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(acc[a][e]);
        }
    }
    if (ok && !(dbg & 8)) {  // (the TMEM loads above are warp-collective: no divergence before them)
      const int col = n0 + col0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (de.bias) bz = __ldg(reinterpret_cast<const float4*>(de.bias + col + 4 * k));
        v[4 * k + 0] = v[4 * k + 0] * de.acc_scale + bz.x;
        v[4 * k + 1] = v[4 * k + 1] * de.acc_scale + bz.y;
        v[4 * k + 2] = v[4 * k + 2] * de.acc_scale + bz.z;
        v[4 * k + 3] = v[4 * k + 3] * de.acc_scale + bz.w;
      }
      if (kind == 1) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] += pre[c][e >> 3].v[e & 7];
      }
      if (gate) {
        float o[8];
#pragma unroll
        for (int p2 = 0; p2 < 8; ++p2) o[p2] = gate_fast(v[2 * p2], v[2 * p2 + 1]) * am * inv_div;
        const int ocol = col >> 1;
        if (de.out) stg256(de.out + (int64_t)b * de.out_bs + (int64_t)row * de.out_ld + ocol, o);
        if (de.out_hi) {
          if (de.out_plane_add) {
            const float4 p0 = __ldg(reinterpret_cast<const float4*>(de.out_plane_add + ocol));
            const float4 p1 = __ldg(reinterpret_cast<const float4*>(de.out_plane_add + ocol + 4));
            o[0] += p0.x; o[1] += p0.y; o[2] += p0.z; o[3] += p0.w;
            o[4] += p1.x; o[5] += p1.y; o[6] += p1.z; o[7] += p1.w;
          }
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int p2 = 0; p2 < 4; ++p2) split2_f16(o[2 * p2], o[2 * p2 + 1], hw[p2], lw[p2]);
          const int64_t pidx = (int64_t)b * de.out_plane_bs + (int64_t)row * de.out_plane_ld + ocol;
          *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(de.out_hi) + pidx) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(de.out_lo) + pidx) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      } else {
        if (de.act == PTTSPP_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
        } else if (de.act != PTTSPP_ACT_NONE) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = act_apply(v[e], de.act);
        }
        if (am != 1.f) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] *= am;
        }
        if (de.res) {
          f8 r0 = pre[c][0], r1 = pre[c][1];
          if (kind != 2) {
            const float* src = de.res + (int64_t)b * de.res_bs + (int64_t)row * de.res_ld + col;
            r0 = ldg256(src);
            r1 = ldg256(src + 8);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[e] += de.res_scale * r0.v[e];
            v[8 + e] += de.res_scale * r1.v[e];
          }
        } else if (kind == 4) {
          // residual from operand planes: r = hi + lo - sub  (pre[c][0] = 16 hi halves, pre[c][1] = 16 lo halves)
#pragma unroll
          for (int p2 = 0; p2 < 8; ++p2) {
            const uint32_t hb = __float_as_uint(pre[c][0].v[p2]), lb = __float_as_uint(pre[c][1].v[p2]);
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hb));
            const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lb));
            float r0 = hf.x + lf.x, r1 = hf.y + lf.y;
            if (de.res_plane_sub) {
              const float2 sb = __ldg(reinterpret_cast<const float2*>(de.res_plane_sub + col + 2 * p2));
              r0 -= sb.x;
              r1 -= sb.y;
            }
            v[2 * p2] += de.res_scale * r0;
            v[2 * p2 + 1] += de.res_scale * r1;
          }
        }
        if (de.out && de.beta != 0.f) {
          f8 r0 = pre[c][0], r1 = pre[c][1];
          if (kind != 3) {
            const float* src = de.out + (int64_t)b * de.out_bs + (int64_t)row * de.out_ld + col;
            r0 = ldg256(src);
            r1 = ldg256(src + 8);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[e] += de.beta * r0.v[e];
            v[8 + e] += de.beta * r1.v[e];
          }
        }
        if (de.out_div != 0.f) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] *= inv_div;
        }
        if (de.out) {
          float* dst = de.out + (int64_t)b * de.out_bs + (int64_t)row * de.out_ld + col;
          stg256(dst, v);
          stg256(dst + 8, v + 8);
        }
        if (de.out_hi) {
          if (de.out_plane_add) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 pz = __ldg(reinterpret_cast<const float4*>(de.out_plane_add + col + 4 * k));
              v[4 * k + 0] += pz.x; v[4 * k + 1] += pz.y; v[4 * k + 2] += pz.z; v[4 * k + 3] += pz.w;
            }
          }
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int p2 = 0; p2 < 8; ++p2) split2_f16(v[2 * p2], v[2 * p2 + 1], hw[p2], lw[p2]);
          const int64_t pidx = (int64_t)b * de.out_plane_bs + (int64_t)row * de.out_plane_ld + col;
          stg256u(reinterpret_cast<__half*>(de.out_hi) + pidx, hw);
          stg256u(reinterpret_cast<__half*>(de.out_lo) + pidx, lw);
        }
      }
    }
    __syncwarp();
  }
}

// Epilogue of one accumulator tile by one of the 16 epilogue warps (TMEM -> registers -> fused epilogue -> HBM).
template <int BN, int NACC>
__device__ __forceinline__ void umma_tile_epilogue(const pttspp_conv1d_desc& de, int n0, int vec_ok, int mt, int b, int i,
                                                   int warp, int lane, uint32_t tmem_base, uint32_t tfull, int n_main,
                                                   const OutMaps* maps, bool tma_st, uint8_t* stage, uint32_t stage_u32) {
  // `de` is one of the kernel's by-value descriptors, selected by a branch at the call site (immediate constant-bank
  // operands instead of indexed LDCs, see umma_tile_epilogue_co)
  const int q = warp & 3;      // TMEM lane quarter this warp may access
  const int cgrp = warp >> 2;  // column group (BN / 4 columns each)
  const int u = i & 1;
  const int m0 = de.m_begin + mt * UM_BM;
  const int m = m0 + q * 32 + lane;
  const int row = m * de.out_mul + de.out_off;
  const bool row_ok = (m < de.m_begin + de.M) && row >= 0 && row < de.T_out;
  float mask = 1.f;
  if (de.out_len && row_ok) mask = ((long long)row < (long long)de.out_len[b]) ? 1.f : 0.f;
  constexpr int CW = BN / (UM_EPI_WARPS / 4);  // columns per warp
  const int cbeg = cgrp * CW;
  const bool vec = vec_ok != 0;
  static_assert(CW == 32 || CW == 16, "one or two 16-column chunks per epilogue warp");
  // operand tiles of BOTH chunks are requested before the accumulator is complete (memory-level parallelism is what
  // bounds the read-modify-write epilogues of the 1x1 projections)
  EpiOps16 ops0, ops1;
  const bool v0 = vec && row_ok && n0 + cbeg + 16 <= de.Cout;
  const bool v1 = vec && row_ok && n0 + cbeg + 32 <= de.Cout;
  if (v0) conv_epilogue16_load(de, b, row, n0 + cbeg, ops0);
  if (CW == 32 && v1) conv_epilogue16_load(de, b, row, n0 + cbeg + 16, ops1);
  mbar_wait_warp(tfull, ((uint32_t)i >> 1) & 1u);
  tc_fence_after();
  const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * NACC * BN);
#pragma unroll
  for (int c = 0; c < CW; c += 16) {
    const int col0 = cbeg + c;
    if (n0 + col0 >= de.Cout) continue;  // warp-uniform: chunk entirely past the last output column
    float v[16], t[16];
    tmem_ld16(tbase + (uint32_t)((NACC - 1) * BN + col0), v);  // cross terms; warp-collective: no divergence before
    for (int a = 0; a < n_main; ++a) {                         // main accumulators (n_main is CTA-uniform)
      tmem_ld16(tbase + (uint32_t)(a * BN + col0), t);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] += t[e];
    }
    if (tma_st) {
      // all 32 rows of the warp leave together: registers -> 2 KB staging tile -> one TMA bulk store per output
      // tensor (rows past the end of the utterance are clipped by the tensor map)
      const bool gate = (de.act == PTTSPP_ACT_GATE);
      const int nout = gate ? 8 : 16;
      const int ocol = gate ? ((n0 + col0) >> 1) : (n0 + col0);
      float o[16];
      conv_epilogue16_math(de, b, row, mask, n0 + col0, v, c == 0 ? ops0 : ops1, o);
      const int row_first = m0 + q * 32;
      if (de.out) {
        if (lane == 0) bulk_wait_read0();  // the previous store has finished reading the staging tile
        __syncwarp();
        float4* st = reinterpret_cast<float4*>(stage + lane * nout * 4);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          if (4 * k4 < nout) st[k4] = make_float4(o[4 * k4], o[4 * k4 + 1], o[4 * k4 + 2], o[4 * k4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&maps->f32, stage_u32, ocol, row_first, b);
          bulk_commit();
        }
      }
      if (de.out_hi) {
        if (de.out_plane_add) {
          const float4* pp = reinterpret_cast<const float4*>(de.out_plane_add + ocol);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            if (4 * k4 < nout) {
              const float4 tq = __ldg(pp + k4);
              o[4 * k4] += tq.x; o[4 * k4 + 1] += tq.y; o[4 * k4 + 2] += tq.z; o[4 * k4 + 3] += tq.w;
            }
        }
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        uint4* sh = reinterpret_cast<uint4*>(stage + lane * nout * 2);
        uint4* sl = reinterpret_cast<uint4*>(stage + 1024 + lane * nout * 2);
#pragma unroll
        for (int k8 = 0; k8 < 2; ++k8)
          if (8 * k8 < nout) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
              __half h0, l0, h1, l1;
              split_f16(o[8 * k8 + 2 * j2], h0, l0);
              split_f16(o[8 * k8 + 2 * j2 + 1], h1, l1);
              hw[j2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
              lw[j2] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            sh[k8] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            sl[k8] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&maps->hi, stage_u32, ocol, row_first, b);
          tma_store_3d(&maps->lo, stage_u32 + 1024, ocol, row_first, b);
          bulk_commit();
        }
      }
    } else if (row_ok) {
      if (c == 0 ? v0 : v1) {
        conv_epilogue16_finish(de, b, row, mask, n0 + col0, v, c == 0 ? ops0 : ops1);
      } else {
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const float a4[4] = {v[gq * 4 + 0], v[gq * 4 + 1], v[gq * 4 + 2], v[gq * 4 + 3]};
          conv_epilogue4(de, b, row, mask, n0 + col0 + gq * 4, a4);
        }
      }
    }
  }
}

// Persistent, warp-specialised: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The TMA and MMA warps
// run ahead into the next tile while the eight epilogue warps drain the previous accumulator: TMEM holds two
// accumulator buffers of (main | cross-term) x BN columns.
// EPI: 0 = TMA bulk-store / direct epilogue, 1 = coalescing epilogue, 2 = CHUNKED near-fp32 accumulation: the K loop is
// cut into chunks of `chunk_iters` (slab, tap) iterations, every chunk accumulates into a FRESH accumulator buffer
// (<= 16 * chunk_iters / 4 truncating tensor-core accumulations each) and the epilogue warps add the chunk results in
// round-to-nearest fp32 registers while the next chunk runs -- the contraction stays in the fp32 error class however
// long K is (the text encoder's k9 feed-forward convs, whose outputs decide integer durations).
// DUAL: the launch carries a second epilogue
// descriptor.  Both are compile-time so that a kernel holds exactly one epilogue body per descriptor (with both
// variants inlined the 96-register budget spilled ~2 KB per thread and the narrow kernel lost 25 %).
template <int BN, int STAGES, int NACC, int EPI, bool DUAL>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv1d_umma_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                   const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                   const __grid_constant__ OutMaps om, const __grid_constant__ OutMaps om2,
                   const pttspp_conv1d_desc d, const pttspp_conv1d_desc d2, const int cout1, const int vec_ok,
                   const int tma_out, const int n_mt, const int n_nt, const int n_tiles, const int chunk_iters) {
  // d describes the contraction (shared by all tiles) and the epilogue of output columns [0, cout1); d2 (dual mode,
  // cout1 < d.Cout) the epilogue of columns [cout1, d.Cout) -- e.g. the residual and skip halves of one projection.
  using SM = UmmaSmem<BN>;
  constexpr uint32_t TMEM_COLS = 2 * NACC * BN;  // 2 buffers x (main accumulators, cross)
  static_assert(TMEM_COLS <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + STAGES * SM::STAGE_BYTES;        // full[S], empty[S], tfull[2], tempty[2]
  constexpr int NBARS = 2 * STAGES + 4;
  const uint32_t tmem_slot = bars + NBARS * 8;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(gen_base + STAGES * SM::STAGE_BYTES + NBARS * 8);
  auto full_bar = [&](int s) { return bars + s * 8; };
  auto empty_bar = [&](int s) { return bars + (STAGES + s) * 8; };
  auto tfull_bar = [&](int u) { return bars + (2 * STAGES + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (2 * STAGES + 2 + u) * 8; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslab = d.Cin / UM_BK;
  const int n_iter = nslab * d.K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), UM_EPI_WARPS);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == UM_EPI_WARPS && lane == 0) {
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapBh);
    tma_prefetch_desc(&mapBl);
  }
  if (warp == UM_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == UM_EPI_WARPS) {
    // ================= TMA producer (whole warp walks the loop, one elected lane issues) =================
    {
      uint32_t g = 0;  // ring position, continues across tiles
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % n_nt, mt = (tile / n_nt) % n_mt, b = tile / (n_nt * n_mt);
        const int m0 = d.m_begin + mt * UM_BM, n0 = nt * BN;
        for (int it = 0; it < n_iter; ++it, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1u;
          mbar_wait_warp(empty_bar(s), ph ^ 1u);
          const int slab = it / d.K, tap = it % d.K;  // taps innermost: the shifted row windows overlap in L2
          const uint32_t st = base + s * SM::STAGE_BYTES;
          const int row = m0 + tap * d.dil - d.pad;
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), SM::STAGE_BYTES);
            tma_load_3d(st, &mapAh, full_bar(s), slab * UM_BK, row, b);
            tma_load_3d(st + SM::A_BYTES, &mapAl, full_bar(s), slab * UM_BK, row, b);
            tma_load_2d(st + 2 * SM::A_BYTES, &mapBh, full_bar(s), slab * UM_BK, tap * d.Cout + n0);
            tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES, &mapBl, full_bar(s), slab * UM_BK, tap * d.Cout + n0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == UM_EPI_WARPS + 1) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
    {
      constexpr uint32_t idesc = umma_idesc_f16(UM_BM, BN);
      const uint64_t desc0 = umma_desc_k_sw128(base);
      uint32_t g = 0;
      int i = 0;
      const int chunk = (EPI == 2) ? chunk_iters : n_iter;  // iterations per accumulator buffer
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int it0 = 0; it0 < n_iter; it0 += chunk, ++i) {
          const int u = i & 1;
          const int it1 = min(n_iter, it0 + chunk);
          mbar_wait_warp(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u);  // epilogue has drained this buffer
          tc_fence_after();
          // The tensor core truncates when it adds into the fp32 accumulator, so the error grows linearly with the
          // number of accumulations into one accumulator: the 2^-11 smaller cross terms (hi*lo, lo*hi) get their
          // own accumulator, the epilogue adds the two in round-to-nearest fp32.
          // NACC - 1 main accumulators are used round-robin by K iteration: the truncation bias of one accumulator
          // grows with the number of MMAs that add into it.
          const uint32_t acc0 = tmem_base + (uint32_t)(u * NACC * BN);
          const uint32_t acc_cross = acc0 + (uint32_t)((NACC - 1) * BN);
          for (int it = it0; it < it1; ++it, ++g) {
            const int s = g % STAGES;
            const uint32_t ph = (g / STAGES) & 1u;
            mbar_wait_warp(full_bar(s), ph);
            tc_fence_after();
            // descriptors of this stage: only the 14-bit start-address field differs from the stage-0 descriptor
            const uint64_t dAh = desc0 + (uint64_t)((uint32_t)s * (uint32_t)(SM::STAGE_BYTES >> 4));
            const uint64_t dAl = dAh + (uint64_t)(SM::A_BYTES >> 4);
            const uint64_t dBh = dAh + (uint64_t)((2 * SM::A_BYTES) >> 4);
            const uint64_t dBl = dAh + (uint64_t)((2 * SM::A_BYTES + SM::B_BYTES) >> 4);
            const int rel = it - it0;
            const uint32_t acc_main = acc0 + (uint32_t)((rel % (NACC - 1)) * BN);
            const uint32_t first_main = (rel >= NACC - 1) ? 1u : 0u;
            if (elect_one()) {
              if constexpr (NACC == 2) {
                // one main accumulator directly followed by the cross accumulator, and the stage holds [W_hi | W_lo] back
                // to back: x_hi . [W_hi | W_lo] is ONE N = 2*BN instruction (main | cross), then x_lo . W_hi -- the A planes
                // are fetched from shared memory twice instead of three times per k step (the operand fetch path, not the
                // tensor pipe, bounds the narrow tiles: ncu tc wavefronts 76 % on the weight-resident kernel)
                constexpr uint32_t idesc2 = umma_idesc_f16(UM_BM, 2 * BN);
#pragma unroll
                for (int kk = 0; kk < UM_BK / 16; ++kk) {
                  const uint64_t adv = (uint64_t)(kk * 2);
                  umma_f16(acc0, dAh + adv, dBh + adv, idesc2, (kk != 0) ? 1u : (rel != 0 ? 1u : 0u));
                  umma_f16(acc_cross, dAl + adv, dBh + adv, idesc, 1u);
                }
              } else {
#pragma unroll
              for (int kk = 0; kk < UM_BK / 16; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 2);  // 16 halves = 32 bytes along K inside the swizzle span
                umma_f16(acc_cross, dAl + adv, dBh + adv, idesc, (kk != 0) ? 1u : (rel != 0 ? 1u : 0u));
                umma_f16(acc_cross, dAh + adv, dBl + adv, idesc, 1u);
                umma_f16(acc_main, dAh + adv, dBh + adv, idesc, (kk != 0) ? 1u : first_main);
              }
              }
              umma_commit(empty_bar(s));  // frees the stage once these MMAs have read it
              if (it + 1 == it1) umma_commit(tfull_bar(u));  // accumulator buffer u complete
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ================= epilogue: warps 0-15 =================
    int i = 0;
    if constexpr (EPI == 2) {
      // chunked accumulation: this warp owns CW columns of 32 rows; running sums live in registers
      constexpr int CW = BN / (UM_EPI_WARPS / 4);
      static_assert(CW == 32, "chunked epilogue: 32 columns per warp");
      const int q = warp & 3, cgrp = warp >> 2;
      const int cbeg = cgrp * CW;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % n_nt, mt = (tile / n_nt) % n_mt, b = tile / (n_nt * n_mt);
        const int n0 = nt * BN;
        const int m = d.m_begin + mt * UM_BM + q * 32 + lane;
        const int row = m * d.out_mul + d.out_off;
        const bool row_ok = (m < d.m_begin + d.M) && row >= 0 && row < d.T_out;
        float mask = 1.f;
        if (d.out_len && row_ok) mask = ((long long)row < (long long)d.out_len[b]) ? 1.f : 0.f;
        EpiOps16 ops0, ops1;
        const bool v0 = vec_ok && row_ok && n0 + cbeg + 16 <= d.Cout;
        const bool v1 = vec_ok && row_ok && n0 + cbeg + 32 <= d.Cout;
        if (v0) conv_epilogue16_load(d, b, row, n0 + cbeg, ops0);
        if (v1) conv_epilogue16_load(d, b, row, n0 + cbeg + 16, ops1);
        float sum[CW];
#pragma unroll
        for (int e = 0; e < CW; ++e) sum[e] = 0.f;
        for (int it0 = 0; it0 < n_iter; it0 += chunk_iters, ++i) {
          const int u = i & 1;
          const int nm = min(min(n_iter - it0, chunk_iters), NACC - 1);
          mbar_wait_warp(tfull_bar(u), ((uint32_t)i >> 1) & 1u);
          tc_fence_after();
          const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * NACC * BN);
#pragma unroll
          for (int c = 0; c < CW; c += 16) {
            float v[16], t[16];
            tmem_ld16(tbase + (uint32_t)((NACC - 1) * BN + cbeg + c), v);  // cross terms
            for (int a = 0; a < nm; ++a) {
              tmem_ld16(tbase + (uint32_t)(a * BN + cbeg + c), t);
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] += t[e];
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) sum[c + e] += v[e];  // round-to-nearest fp32 across chunks
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(u));
        }
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < CW; c += 16) {
            const int col0 = cbeg + c;
            if (n0 + col0 >= d.Cout) continue;
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = sum[c + e];
            if (c == 0 ? v0 : v1) {
              conv_epilogue16_finish(d, b, row, mask, n0 + col0, v, c == 0 ? ops0 : ops1);
            } else {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                const float a4[4] = {v[gq * 4 + 0], v[gq * 4 + 1], v[gq * 4 + 2], v[gq * 4 + 3]};
                conv_epilogue4(d, b, row, mask, n0 + col0 + gq * 4, a4);
              }
            }
          }
        }
      }
    } else
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
      const int nt = tile % n_nt, mt = (tile / n_nt) % n_mt, b = tile / (n_nt * n_mt);
      const int nm = n_iter < NACC - 1 ? n_iter : NACC - 1;
      const bool second = DUAL && nt * BN >= cout1;  // CTA-uniform branch: each side reads its descriptor with immediates
      if constexpr (EPI == 1) {
        float* stg = reinterpret_cast<float*>(gen_base + STAGES * SM::STAGE_BYTES + 256 + warp * 2048);
        if (second)
          umma_tile_epilogue_co<BN, NACC, UM_EPI_WARPS>(d2, nt * BN - cout1, mt, b, i & 1, ((uint32_t)i >> 1) & 1u, warp,
                                                        lane, tmem_base, tfull_bar(i & 1), nm, stg, 0);
        else
          umma_tile_epilogue_co<BN, NACC, UM_EPI_WARPS>(d, nt * BN, mt, b, i & 1, ((uint32_t)i >> 1) & 1u, warp, lane,
                                                        tmem_base, tfull_bar(i & 1), nm, stg, 0);
      } else {
        uint8_t* stg = gen_base + STAGES * SM::STAGE_BYTES + 256 + warp * 2048;
        const uint32_t stg_u32 = base + STAGES * SM::STAGE_BYTES + 256 + warp * 2048;
        if (second)
          umma_tile_epilogue<BN, NACC>(d2, nt * BN - cout1, vec_ok, mt, b, i, warp, lane, tmem_base, tfull_bar(i & 1), nm,
                                       &om2, (tma_out & 2) != 0, stg, stg_u32);
        else
          umma_tile_epilogue<BN, NACC>(d, nt * BN, vec_ok, mt, b, i, warp, lane, tmem_base, tfull_bar(i & 1), nm, &om,
                                       (tma_out & 1) != 0, stg, stg_u32);
      }
      const int u = i & 1;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(u));
    }
    if ((tma_out & 3) && lane == 0) bulk_wait0();  // all bulk stores of this warp have been written
  }
  tc_fence_before();
  __syncthreads();
  if (warp == UM_EPI_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ---- A-stationary variant ---------------------------------------------------------------------------------------
// Shared-memory fill bandwidth (L2 -> SM) bounds the streaming kernel above: every (tap, N tile) re-loads its
// activation slab.  Here a CTA owns 128 output rows of one utterance, loads their activation planes ONCE -- all Cin
// slabs, 128 + halo rows -- and walks all N tiles and taps over them, streaming only weights through the ring.
// Taps share the halo tile: the UMMA descriptor start is simply advanced by tap*dil rows (the swizzle XOR is taken
// from absolute shared-memory address bits, so any row offset is legal -- verified on B200 by tools/probe_desc.py).
// NSUB = 2: a unit is 256 rows = two 128-row MMA tiles that share every weight stage (the weight tiles are what all SMs
// re-stream from the same L2 lines: ncu of the 64-column k = 11 conv showed 5.4 TB/s of L2 -> SM fill at 36 % tensor pipe).
template <int BN, int NSUB, bool RL>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv1d_umma_as_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                      const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                      const pttspp_conv1d_desc d, const pttspp_conv1d_desc d2, const int cout1, const int cout_total,
                      const int vec_ok, const int n_mt, const int n_nt, const int n_units, const int rowsA,
                      const int nbst, const int nabuf) {
  constexpr uint32_t TMEM_COLS = 2 * NSUB * UM_NACC * BN;
  static_assert(TMEM_COLS <= 512, "TMEM budget");
  constexpr int B_BYTES = BN * 128;        // one plane of one weight tile
  constexpr int BST_BYTES = 2 * B_BYTES;   // hi + lo
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nslab = d.Cin / UM_BK;
  const uint32_t a_plane = (uint32_t)rowsA * 128u;          // bytes of one (slab, plane) tile; rowsA % 8 == 0
  const uint32_t a_bytes = (uint32_t)nslab * 2u * a_plane;  // whole activation block
  // nabuf = 2: the next unit's activation block streams in while this unit's MMAs run (a must when a unit is a single
  // N tile, e.g. BigVGAN's 64-channel k = 11 convs: otherwise block load and MMAs alternate)
  const uint32_t ring = base + (uint32_t)nabuf * a_bytes;
  const uint32_t bars = ring + (uint32_t)nbst * BST_BYTES;  // fullA[2], emptyA[2], fullB[nbst], emptyB[nbst], tfull[2], tempty[2]
  const int nbars = 4 + 2 * nbst + 4;
  const uint32_t tmem_slot = bars + nbars * 8;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + nabuf * a_bytes + nbst * BST_BYTES + nbars * 8);
  auto fullA = [&](int q) { return bars + q * 8; };
  auto emptyA = [&](int q) { return bars + (2 + q) * 8; };
  auto fullB = [&](int s) { return bars + (4 + s) * 8; };
  auto emptyB = [&](int s) { return bars + (4 + nbst + s) * 8; };
  auto tfull_bar = [&](int u) { return bars + (4 + 2 * nbst + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (4 + 2 * nbst + 2 + u) * 8; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 2; ++q) {
      mbar_init(fullA(q), 1);
      mbar_init(emptyA(q), 1);
    }
    for (int s = 0; s < nbst; ++s) {
      mbar_init(fullB(s), 1);
      mbar_init(emptyB(s), 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), UM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == UM_EPI_WARPS && lane == 0) {
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapBh);
    tma_prefetch_desc(&mapBl);
  }
  if (warp == UM_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == UM_EPI_WARPS) {
    // ================= TMA producer (warp-uniform loop, one elected lane issues) =================
    uint32_t g = 0;
    int j = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++j) {
      const int mt = unit % n_mt, b = unit / n_mt;  // n_mt counts units (NSUB x 128 rows) per utterance
      const int row0 = d.m_begin + mt * (NSUB * UM_BM) - d.pad;  // first input row of the halo block (may be negative)
      const int nbox = (rowsA > 256) ? 2 : 1, box_rows = rowsA / nbox;  // a TMA box holds at most 256 rows
      const int q = j % nabuf;
      const uint32_t ablk = base + (uint32_t)q * a_bytes;
      mbar_wait_warp(emptyA(q), (((uint32_t)(j / nabuf)) & 1u) ^ 1u);  // the MMAs of the unit that used this buffer have retired
      if (elect_one()) {
        mbar_expect_tx(fullA(q), a_bytes);
        for (int slab = 0; slab < nslab; ++slab)
          for (int bx = 0; bx < nbox; ++bx) {
            const uint32_t boff = (uint32_t)(bx * box_rows) * 128u;
            tma_load_3d(ablk + (uint32_t)(2 * slab) * a_plane + boff, &mapAh, fullA(q), slab * UM_BK, row0 + bx * box_rows, b);
            tma_load_3d(ablk + (uint32_t)(2 * slab + 1) * a_plane + boff, &mapAl, fullA(q), slab * UM_BK, row0 + bx * box_rows, b);
          }
      }
      __syncwarp();
      for (int nt = 0; nt < n_nt; ++nt)
        for (int slab = 0; slab < nslab; ++slab)
          for (int tap = 0; tap < d.K; ++tap, ++g) {
            const int s = g % nbst;
            const uint32_t ph = (g / nbst) & 1u;
            mbar_wait_warp(emptyB(s), ph ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(fullB(s), BST_BYTES);
              const uint32_t st = ring + (uint32_t)s * BST_BYTES;
              tma_load_2d(st, &mapBh, fullB(s), slab * UM_BK, tap * cout_total + nt * BN);
              tma_load_2d(st + B_BYTES, &mapBl, fullB(s), slab * UM_BK, tap * cout_total + nt * BN);
            }
            __syncwarp();
          }
    }
  } else if (warp == UM_EPI_WARPS + 1) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues: barrier addresses and descriptors stay in
    // uniform registers; the single-lane form of this loop issued one instruction every few cycles and was the
    // kernel's bottleneck -- ncu: no barrier retries in this warp at 36 % tensor pipe) =================
    constexpr uint32_t idesc = umma_idesc_f16(UM_BM, BN);
    constexpr uint32_t idesc2 = umma_idesc_f16(UM_BM, 2 * BN);  // [W_hi | W_lo] stacked along N (main | cross)
    const uint64_t desc0 = umma_desc_k_sw128(base);  // descriptors differ only in the 14-bit start-address field
    uint32_t g = 0;
    int j = 0, i = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++j) {
      const int q = j % nabuf;
      const uint32_t ablk_off = (uint32_t)q * a_bytes;
      mbar_wait_warp(fullA(q), ((uint32_t)(j / nabuf)) & 1u);
      tc_fence_after();
      for (int nt = 0; nt < n_nt; ++nt, ++i) {
        const int u = i & 1;
        mbar_wait_warp(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t acc_u = tmem_base + (uint32_t)(u * NSUB * UM_NACC * BN);  // sub-tile s: + s * NACC * BN (main | cross)
        uint32_t first = 0;
        for (int slab = 0; slab < nslab; ++slab)
          for (int tap = 0; tap < d.K; ++tap, ++g) {
            const int s = g % nbst;
            const uint32_t ph = (g / nbst) & 1u;
            mbar_wait_warp(fullB(s), ph);
            tc_fence_after();
            const uint32_t a_off = ablk_off + (uint32_t)(2 * slab) * a_plane + (uint32_t)(tap * d.dil) * 128u;  // taps share the halo tile
            const uint64_t dAh = desc0 + (uint64_t)(a_off >> 4);
            const uint64_t dAl = dAh + (uint64_t)(a_plane >> 4);
            const uint64_t dBh = desc0 + (uint64_t)(((uint32_t)nabuf * a_bytes + (uint32_t)s * BST_BYTES) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub) {
                const uint64_t soff = (uint64_t)((uint32_t)(sub * UM_BM) * 128u >> 4);  // the sub-tile's rows of the block
                const uint32_t acc_main = acc_u + (uint32_t)(sub * UM_NACC * BN), acc_cross = acc_main + (uint32_t)BN;
#pragma unroll
                for (int kk = 0; kk < UM_BK / 16; ++kk) {
                  const uint64_t adv = (uint64_t)(kk * 32 >> 4) + soff;
                  const uint64_t advb = (uint64_t)(kk * 32 >> 4);
                  // the streaming kernel's stacked issue, same order: the two kernels give the same bits
                  umma_f16(acc_main, dAh + adv, dBh + advb, idesc2, (kk != 0) ? 1u : first);
                  umma_f16(acc_cross, dAl + adv, dBh + advb, idesc, 1u);
                }
              }
              umma_commit(emptyB(s));
            }
            __syncwarp();
            first = 1u;
          }
        if (elect_one()) umma_commit(tfull_bar(u));
        __syncwarp();
      }
      if (elect_one()) umma_commit(emptyA(q));  // the activation block may be overwritten once everything issued so far has retired
      __syncwarp();
    }
  } else {
    // ================= epilogue: warps 0-15 =================
    int i = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int mt = unit % n_mt, b = unit / n_mt;
      for (int nt = 0; nt < n_nt; ++nt, ++i) {
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          // buffer (i & 1), sub-tile `sub`: the epilogue addresses buffer (its i) & 1 and waits with parity (its i >> 1) & 1
          const uint32_t tb = tmem_base + (uint32_t)(((i & 1) * NSUB + sub) * UM_NACC * BN);
          const int i2 = i & ~1;
          if constexpr (RL) {  // row-per-lane 256-bit epilogue (host: epilogue_rl_ok): half the instructions of the generic one
            if (nt * BN >= cout1)
              umma_tile_epilogue_rl<BN, UM_NACC, UM_EPI_WARPS>(d2, nt * BN - cout1, mt * NSUB + sub, b, 0, ((uint32_t)i >> 1) & 1u,
                                                               warp, lane, tb, tfull_bar(i & 1), 1, 0);
            else
              umma_tile_epilogue_rl<BN, UM_NACC, UM_EPI_WARPS>(d, nt * BN, mt * NSUB + sub, b, 0, ((uint32_t)i >> 1) & 1u, warp,
                                                               lane, tb, tfull_bar(i & 1), 1, 0);
          } else if (nt * BN >= cout1)
            umma_tile_epilogue<BN, UM_NACC>(d2, nt * BN - cout1, vec_ok, mt * NSUB + sub, b, i2, warp, lane, tb,
                                            tfull_bar(i & 1), 1, nullptr, false, nullptr, 0u);
          else
            umma_tile_epilogue<BN, UM_NACC>(d, nt * BN, vec_ok, mt * NSUB + sub, b, i2, warp, lane, tb, tfull_bar(i & 1), 1,
                                            nullptr, false, nullptr, 0u);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(i & 1));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == UM_EPI_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ---- CTA-pair (cta_group::2), A-stationary kernel ------------------------------------------------------------------
// The streaming kernel above is bound by the L2 -> SM fill rate (ncu: 11.3 TB/s = the LTS cap, tensor pipe 45 %):
// every (tap, N tile) re-loads its activation slab and every CTA loads the whole weight tile.  Here two CTAs of one
// TPC form a pair: one tcgen05.mma.cta_group::2 computes 256 rows x 128 columns, each CTA supplying ITS 128
// activation rows and HALF (64 rows) of the weight tile -- weight traffic per SM is halved.  The activation block of
// a 256-row unit (all Cin slabs, 128 + halo rows per CTA) is loaded ONCE per unit and all N tiles / taps walk over
// it (taps share the halo tile by advancing the descriptor start by tap*dil rows); only the 16 KB weight half-tiles
// stream through a ring.  Per-slab full/empty barriers let the next unit's block stream in while the last tile of
// the current unit still computes.  Tiles (unit, N tile) are dealt to the clusters in contiguous chunks, so the
// machine is filled to 1/tiles granularity instead of 1/units.
// Barrier ownership: "full" barriers live in the leader (even) CTA and receive the TMA bytes of both CTAs;
// "empty"/"tfull" barriers exist in both CTAs and are signalled by multicast tcgen05.commit; the leader's "tempty"
// collects the epilogue warps of both CTAs (remote arrive).
constexpr int UP_BN = 128;
constexpr int UP_BHALF = UP_BN / 2;
constexpr int UP_BST_BYTES = 2 * UP_BHALF * 128;  // hi + lo half-tiles of one stage
constexpr int UP_MAX_SLAB = 8;

// EW epilogue warps (4 TMEM lane quarters x EW/4 column groups): 16 drain a tile fastest, 8 leave shared memory for
// one more weight stage.
// NACC = 2: main | cross-term accumulators, two tile buffers.  NACC = 1 (K*Cin <= 768: at most 144 accumulations,
// truncation bias ~1.5e-5 relative): one accumulator per tile and FOUR tile buffers, so the MMA warp runs up to three
// tiles ahead of a memory-bound epilogue.
template <int NBST, int EW, int NACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((EW + 2) * 32, 1)
conv1d_umma_pair_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                        const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                        const pttspp_conv1d_desc d, const pttspp_conv1d_desc d2, const int cout1, const int cout_total,
                        const int n_mt, const int n_nt, const int n_tiles, const int rowsA, const int mma_order,
                        const int epi_rl) {
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int NBUF = 512 / (NACC * UP_BN);  // accumulator tile buffers
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int nslab = d.Cin / UM_BK;
  const uint32_t a_plane = (uint32_t)rowsA * 128u;
  const uint32_t a_bytes = (uint32_t)nslab * 2u * a_plane;
  const uint32_t ring = base + a_bytes;
  const uint32_t stg_off = a_bytes + (uint32_t)NBST * UP_BST_BYTES;  // epilogue staging, 2 KB per warp
  const uint32_t bars_off = stg_off + (epi_rl ? 0u : (uint32_t)EW * 2048u);  // the row-per-lane gate epilogue needs no staging
  const uint32_t bars = base + bars_off;
  // fullA[8], emptyA[8], fullB[NBST], emptyB[NBST], tfull[2], tempty[2]
  constexpr int NBARS = 2 * UP_MAX_SLAB + 2 * NBST + 8;
  auto fullA = [&](int sl) { return bars + sl * 8; };
  auto emptyA = [&](int sl) { return bars + (UP_MAX_SLAB + sl) * 8; };
  auto fullB = [&](int st) { return bars + (2 * UP_MAX_SLAB + st) * 8; };
  auto emptyB = [&](int st) { return bars + (2 * UP_MAX_SLAB + NBST + st) * 8; };
  auto tfull_bar = [&](int u) { return bars + (2 * UP_MAX_SLAB + 2 * NBST + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (2 * UP_MAX_SLAB + 2 * NBST + 4 + u) * 8; };
  const uint32_t tmem_slot = bars + NBARS * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bars_off + NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int t_begin = (int)((long long)cluster_id * n_tiles / n_clusters);
  const int t_end = (int)((long long)(cluster_id + 1) * n_tiles / n_clusters);

  if (threadIdx.x == 0) {
    for (int sl = 0; sl < UP_MAX_SLAB; ++sl) {
      mbar_init(fullA(sl), 1);
      mbar_init(emptyA(sl), 1);
    }
    for (int st = 0; st < NBST; ++st) {
      mbar_init(fullB(st), 1);
      mbar_init(emptyB(st), 1);
    }
    for (int u = 0; u < NBUF; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), 2 * EW);  // epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == EW && lane == 0) {
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapBh);
    tma_prefetch_desc(&mapBl);
  }
  if (warp == EW + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and the TMEM allocation of both CTAs are visible pair-wide
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == EW) {
    // ================= TMA producer (both CTAs: own activation rows, own half of every weight tile) =================
    // the whole warp walks the loop (warp-uniform control flow), one elected lane issues
    {
      uint32_t g = 0;
      int ablk = 0, cur_unit = -1;
      for (int t = t_begin; t < t_end; ++t) {
        const int unit = t / n_nt, nt = t - unit * n_nt;
        const bool new_unit = unit != cur_unit;
        cur_unit = unit;
        // a unit is TWO independent 128-row blocks (one per CTA) of the flat (utterance, block) list: the pair's CTAs
        // may work on different utterances, so only the very last unit can be half empty
        const int q = 2 * unit + (int)rank;
        const int b = min(q / n_mt, d.B - 1), mt = (q / n_mt < d.B) ? q - (q / n_mt) * n_mt : n_mt;  // past the end: no rows
        const int row0 = d.m_begin + mt * UM_BM - d.pad;
        if (mma_order & 128) {  // experiment (measured harmful: the bulk prefetches queue in front of the TMA loads)
          // The producer runs one to two tiles ahead of the epilogue: pull the epilogue's operand tile (conditioner /
          // residual / previous output, 128 rows x 512 B) into L2 now, so that the epilogue's loads are L2 hits.
          const bool second = nt * UP_BN >= cout1;
          const pttspp_conv1d_desc& de = second ? d2 : d;
          const int n0 = second ? nt * UP_BN - cout1 : nt * UP_BN;
          const int kind = conv_epilogue_prefetch_kind(de);
          if (kind != 0 && n0 + UP_BN <= de.Cout) {
            const bool gate = (de.act == PTTSPP_ACT_GATE);
            const float* src = (kind == 1) ? de.addend : (kind == 2 ? de.res : de.out);
            const int64_t bs = (kind == 1) ? de.addend_bs : (kind == 2 ? de.res_bs : de.out_bs);
            const int ld = (kind == 1) ? de.addend_ld : (kind == 2 ? de.res_ld : de.out_ld);
            const int c0 = (kind == 1 || !gate) ? n0 : (n0 >> 1);
            const uint32_t bytes = (kind == 1 || !gate) ? UP_BN * 4 : UP_BN * 2;
#pragma unroll
            for (int rr = 0; rr < UM_BM / 32; ++rr) {
              const int m = d.m_begin + mt * UM_BM + rr * 32 + lane;
              const int row = m * de.out_mul + de.out_off;
              if (m < de.m_begin + de.M && row >= 0 && row < de.T_out)
                l2_prefetch_bulk(src + (int64_t)b * bs + (int64_t)row * ld + c0, bytes);
            }
          }
          __syncwarp();
        }
        for (int slab = 0; slab < nslab; ++slab) {
          if (new_unit) {
            // this slab of the previous unit has been consumed (released early in that unit's last tile)
            mbar_wait_warp(emptyA(slab), ((uint32_t)ablk & 1u) ^ 1u);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(fullA(slab), 4u * a_plane);
              tma_load_3d_pair(base + (uint32_t)(2 * slab) * a_plane, &mapAh, fullA(slab), slab * UM_BK, row0, b);
              tma_load_3d_pair(base + (uint32_t)(2 * slab + 1) * a_plane, &mapAl, fullA(slab), slab * UM_BK, row0, b);
            }
            __syncwarp();
          }
          for (int tap = 0; tap < d.K; ++tap, ++g) {
            const int st = g % NBST;
            const uint32_t ph = (g / NBST) & 1u;
            mbar_wait_warp(emptyB(st), ph ^ 1u);
            const uint32_t dst = ring + (uint32_t)st * UP_BST_BYTES;
            const int wrow = tap * cout_total + nt * UP_BN + (int)rank * UP_BHALF;
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(fullB(st), 2u * UP_BST_BYTES);
              tma_load_2d_pair(dst, &mapBh, fullB(st), slab * UM_BK, wrow);
              tma_load_2d_pair(dst + UP_BHALF * 128, &mapBl, fullB(st), slab * UM_BK, wrow);
            }
            __syncwarp();
          }
        }
        if (new_unit) ++ablk;
      }
    }
  } else if (warp == EW + 1) {
    // ================= MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) =================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * UM_BM, UP_BN);
      const uint64_t desc0 = umma_desc_k_sw128(base);  // descriptors differ only in the start-address field
      uint32_t g = 0;
      int i = 0, ablk = 0, cur_unit = -1;
      for (int t = t_begin; t < t_end; ++t, ++i) {
        const int unit = t / n_nt;
        const bool new_unit = unit != cur_unit;
        cur_unit = unit;
        const bool last_of_unit = (t + 1 == t_end) || ((t + 1) / n_nt != unit);
        const int u = i % NBUF;
        mbar_wait_warp(tempty_bar(u), (((uint32_t)i / NBUF) & 1u) ^ 1u);  // both CTAs' epilogues have drained buffer u
        tc_fence_after();
        const uint32_t acc_main = tmem_base + (uint32_t)(u * NACC * UP_BN);
        const uint32_t acc_cross = acc_main + (uint32_t)((NACC - 1) * UP_BN);  // NACC == 1: the same accumulator
        uint32_t first = 0;
        for (int slab = 0; slab < nslab; ++slab) {
          if (new_unit) {
            mbar_wait_warp(fullA(slab), (uint32_t)ablk & 1u);
            tc_fence_after();
          }
          for (int tap = 0; tap < d.K; ++tap, ++g) {
            const int st = g % NBST;
            const uint32_t ph = (g / NBST) & 1u;
            mbar_wait_warp(fullB(st), ph);
            tc_fence_after();
            const uint32_t a_off = (uint32_t)(2 * slab) * a_plane + (uint32_t)(tap * d.dil) * 128u;  // taps share the halo tile
            const uint64_t dAh = desc0 + (uint64_t)(a_off >> 4);
            const uint64_t dAl = dAh + (uint64_t)(a_plane >> 4);
            const uint64_t dBh = desc0 + (uint64_t)((a_bytes + (uint32_t)st * UP_BST_BYTES) >> 4);
            const uint64_t dBl = dBh + (uint64_t)((UP_BHALF * 128) >> 4);
            const bool release_a = last_of_unit && (tap + 1 == d.K);
            const bool tile_done = (slab + 1 == nslab) && (tap + 1 == d.K);
            if (elect_one()) {
              if ((mma_order & 3) == 0) {
#pragma unroll
                for (int kk = 0; kk < UM_BK / 16; ++kk) {
                  const uint64_t adv = (uint64_t)(kk * 2);
                  // cross terms in the order of the streaming kernel's stacked issue (hi.lo, then lo.hi): a conv gives
                  // the same bits whichever of the two kernels its batch size selects
                  umma_f16_pair(acc_cross, dAh + adv, dBl + adv, idesc, kk ? 1u : first);
                  umma_f16_pair(acc_cross, dAl + adv, dBh + adv, idesc, 1u);
                  umma_f16_pair(acc_main, dAh + adv, dBh + adv, idesc, (kk || NACC == 1) ? 1u : first);
                }
              } else if ((mma_order & 3) == 1) {
                // grouped by accumulator: the destination changes twice per stage instead of after every MMA
#pragma unroll
                for (int kk = 0; kk < UM_BK / 16; ++kk)
                  umma_f16_pair(acc_cross, dAl + (uint64_t)(kk * 2), dBh + (uint64_t)(kk * 2), idesc, kk ? 1u : first);
#pragma unroll
                for (int kk = 0; kk < UM_BK / 16; ++kk)
                  umma_f16_pair(acc_cross, dAh + (uint64_t)(kk * 2), dBl + (uint64_t)(kk * 2), idesc, 1u);
#pragma unroll
                for (int kk = 0; kk < UM_BK / 16; ++kk)
                  umma_f16_pair(acc_main, dAh + (uint64_t)(kk * 2), dBh + (uint64_t)(kk * 2), idesc, (kk || NACC == 1) ? 1u : first);
              } else {
                // timing experiment only (wrong numerics budget): one accumulator, one MMA per product
#pragma unroll
                for (int kk = 0; kk < UM_BK / 16; ++kk)
                  umma_f16_pair(acc_main, dAh + (uint64_t)(kk * 2), dBh + (uint64_t)(kk * 2), idesc, kk ? 1u : first);
                if (first == 0) umma_f16_pair(acc_cross, dAl, dBh, idesc, 0u);
              }
              umma_commit_pair(emptyB(st));
              if (release_a) umma_commit_pair(emptyA(slab));  // this slab may be overwritten by the next unit
              if (tile_done) umma_commit_pair(tfull_bar(u));
            }
            __syncwarp();
            first = 1u;
          }
        }
        if (new_unit) ++ablk;
      }
    }
  } else {
    // ================= epilogue: warps 0-15 of both CTAs, each CTA drains its own 128 TMEM lanes =================
    int i = 0;
    float* stage = reinterpret_cast<float*>(gen_base + stg_off + warp * 2048);
    for (int t = t_begin; t < t_end; ++t, ++i) {
      const int unit = t / n_nt, nt = t - unit * n_nt;
      const int q = 2 * unit + (int)rank;
      const int b = min(q / n_mt, d.B - 1), mt = (q / n_mt < d.B) ? q - (q / n_mt) * n_mt : n_mt;
      const int ub = i % NBUF;
      const uint32_t tpar = ((uint32_t)i / NBUF) & 1u;
      // CTA-uniform branches: each call reads ITS descriptor(s) with immediate constant operands
      if (epi_rl) {
        if (t + 1 < t_end && !(mma_order & 16)) {
          const int tn = t + 1, unit_n = tn / n_nt, nt_n = tn - unit_n * n_nt;
          const int q_n = 2 * unit_n + (int)rank;
          const int b_n = min(q_n / n_mt, d.B - 1), mt_n = (q_n / n_mt < d.B) ? q_n - (q_n / n_mt) * n_mt : n_mt;
          if (nt_n * UP_BN >= cout1) epi_rl_prefetch_tile<UP_BN, EW>(d2, nt_n * UP_BN - cout1, mt_n, b_n, warp, lane);
          else epi_rl_prefetch_tile<UP_BN, EW>(d, nt_n * UP_BN, mt_n, b_n, warp, lane);
        }
        if (nt * UP_BN >= cout1)
          umma_tile_epilogue_rl<UP_BN, NACC, EW>(d2, nt * UP_BN - cout1, mt, b, ub, tpar, warp, lane, tmem_base, tfull_bar(ub),
                                                 NACC - 1, mma_order);
        else
          umma_tile_epilogue_rl<UP_BN, NACC, EW>(d, nt * UP_BN, mt, b, ub, tpar, warp, lane, tmem_base, tfull_bar(ub), NACC - 1,
                                                 mma_order);
      } else {
        if (nt * UP_BN >= cout1)
          umma_tile_epilogue_co<UP_BN, NACC, EW>(d2, nt * UP_BN - cout1, mt, b, ub, tpar, warp, lane, tmem_base, tfull_bar(ub),
                                                 NACC - 1, stage, mma_order);
        else
          umma_tile_epilogue_co<UP_BN, NACC, EW>(d, nt * UP_BN, mt, b, ub, tpar, warp, lane, tmem_base, tfull_bar(ub), NACC - 1,
                                                 stage, mma_order);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar(i % NBUF));
        else mbar_arrive_cluster(tempty_bar(i % NBUF), 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == EW + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ---- narrow-channel (Cin = Cout = 32) weight-resident kernel -------------------------------------------------------
// BigVGAN's last stage (32 channels, 245 760 samples per utterance) is HBM bound: 1 GB of activation traffic per conv
// against 8-88 GFLOP.  All taps' weight planes (K x 32 x 32 halves x 2 planes <= 44 KB) stay resident in shared memory
// for the whole kernel; per 128-row tile the producer loads ONE halo block (128 + (K-1)*dil rows x 64 B x 2 planes,
// SWIZZLE_64B K-major) and every tap's MMAs read it through a descriptor advanced by tap*dil rows.  The tensor work per
// tile is tiny (K x 6 MMAs of 128 x 32 x 16), so the kernel is organised for memory-level parallelism: a deep
// activation ring, every TMEM column as accumulator buffers, and TWO epilogue groups of 8 warps that alternate tiles, so that one
// group's residual loads are in flight while the other computes and stores.
constexpr int US_EW = 8;            // epilogue warps per group
constexpr int US_GROUPS = 2;
constexpr int US_THREADS = (US_GROUPS * US_EW + 2) * 32;
constexpr int US_MAX_NBUF = 8;       // accumulator buffers (main | cross): 512 TMEM columns / (2 * channels)
constexpr int US_MAXK = 11;

// K-major, SWIZZLE_64B shared-memory matrix descriptor: rows of 64 bytes, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;  // SWIZZLE_64B
  return d;
}

// US_C = 32: 64-byte rows, SWIZZLE_64B; US_C = 64: 128-byte rows, SWIZZLE_128B (BigVGAN's 64-channel stage, taps <= 7:
// K x 64 x 64 x 2 planes <= 112 KB resident).
template <int US_C, int NST>
__global__ void __launch_bounds__(US_THREADS, 1)
conv1d_umma_c32_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                       const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                       const pttspp_conv1d_desc d, const int n_mt, const int n_tiles, const int rowsA) {
  constexpr int NBUF = 512 / (2 * US_C);  // accumulator buffers: all 512 TMEM columns (8 at 32 channels, 4 at 64)
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t ROWB = US_C * 2;                 // bytes per operand row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_plane = (uint32_t)d.K * US_C * ROWB;         // one weight plane: K taps x C rows x ROWB
  const uint32_t w_bytes = 2u * w_plane;
  const uint32_t a_plane = (uint32_t)rowsA * ROWB;
  const uint32_t a_stage = ((2u * a_plane) + 1023u) & ~1023u;
  const uint32_t ring = base + ((w_bytes + 1023u) & ~1023u);
  const uint32_t bars_off = ((w_bytes + 1023u) & ~1023u) + (uint32_t)NST * a_stage;
  const uint32_t bars = base + bars_off;
  // fullW, fullA[NST], emptyA[NST], tfull[NBUF], tempty[NBUF]
  constexpr int NBARS = 1 + 2 * NST + 2 * NBUF;
  const uint32_t fullW = bars;
  auto fullA = [&](int st) { return bars + (1 + st) * 8; };
  auto emptyA = [&](int st) { return bars + (1 + NST + st) * 8; };
  auto tfull_bar = [&](int u) { return bars + (1 + 2 * NST + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (1 + 2 * NST + NBUF + u) * 8; };
  const uint32_t tmem_slot = bars + NBARS * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bars_off + NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_PROD = US_GROUPS * US_EW, W_MMA = W_PROD + 1;
  if (threadIdx.x == 0) {
    mbar_init(fullW, 1);
    for (int st = 0; st < NST; ++st) {
      mbar_init(fullA(st), 1);
      mbar_init(emptyA(st), 1);
    }
    for (int u = 0; u < NBUF; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), US_EW);
    }
    fence_barrier_init();
  }
  if (warp == W_PROD && lane == 0) {
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapBh);
    tma_prefetch_desc(&mapBl);
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == W_PROD) {
    // ================= TMA producer =================
    if (elect_one()) {
      mbar_expect_tx(fullW, w_bytes);
      for (int tap = 0; tap < d.K; ++tap) {  // per tap: the hi rows directly followed by the lo rows (one N = 2C operand)
        tma_load_2d(base + (uint32_t)tap * (2 * US_C * ROWB), &mapBh, fullW, 0, tap * US_C);
        tma_load_2d(base + (uint32_t)tap * (2 * US_C * ROWB) + US_C * ROWB, &mapBl, fullW, 0, tap * US_C);
      }
    }
    __syncwarp();
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
      const int mt = tile % n_mt, b = tile / n_mt;
      const int st = g % NST;
      mbar_wait_warp(emptyA(st), ((g / NST) & 1u) ^ 1u);
      const int row0 = d.m_begin + mt * UM_BM - d.pad;
      if (elect_one()) {
        mbar_expect_tx(fullA(st), 2u * a_plane);
        tma_load_3d(ring + (uint32_t)st * a_stage, &mapAh, fullA(st), 0, row0, b);
        tma_load_3d(ring + (uint32_t)st * a_stage + a_plane, &mapAl, fullA(st), 0, row0, b);
      }
      __syncwarp();
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_f16(UM_BM, US_C);
    constexpr uint32_t idesc2 = umma_idesc_f16(UM_BM, 2 * US_C);  // [W_hi | W_lo] stacked along N: main | cross accumulators
    const uint64_t descW = (US_C == 32) ? umma_desc_k_sw64(base) : umma_desc_k_sw128(base);
    mbar_wait_warp(fullW, 0);
    tc_fence_after();
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
      const int st = g % NST, u = g % NBUF;
      mbar_wait_warp(tempty_bar(u), ((g / NBUF) & 1u) ^ 1u);
      mbar_wait_warp(fullA(st), (g / NST) & 1u);
      tc_fence_after();
      const uint32_t acc_main = tmem_base + (uint32_t)(u * 2 * US_C);
      const uint32_t acc_cross = acc_main + (uint32_t)US_C;
      const uint64_t descA = (US_C == 32) ? umma_desc_k_sw64(ring + (uint32_t)st * a_stage)
                                          : umma_desc_k_sw128(ring + (uint32_t)st * a_stage);
      if (elect_one()) {
        for (int tap = 0; tap < d.K; ++tap) {
          const uint64_t dAh = descA + (uint64_t)(((uint32_t)(tap * d.dil) * ROWB) >> 4);  // taps share the halo block
          const uint64_t dAl = dAh + (uint64_t)(a_plane >> 4);
          const uint64_t dBh = descW + (uint64_t)(((uint32_t)tap * (2 * US_C * ROWB)) >> 4);
#pragma unroll
          for (int kk = 0; kk < US_C / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);  // 16 halves = 32 bytes inside the swizzle span
            const uint32_t acc = (tap | kk) ? 1u : 0u;
            // the kernel is bound by the tensor core's shared-memory operand reads (ncu: tc wavefronts 76 % of peak, a
            // 128-row A tile per instruction): x_hi . [W_hi | W_lo] as ONE N = 2C instruction (main | cross columns are
            // adjacent) + x_lo . W_hi -- the A planes are read twice instead of three times per k step
            umma_f16(acc_main, dAh + adv, dBh + adv, idesc2, acc);
            umma_f16(acc_cross, dAl + adv, dBh + adv, idesc, 1u);
          }
        }
        umma_commit(emptyA(st));
        umma_commit(tfull_bar(u));
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: two groups of 8 warps alternate tiles =================
    const int grp = warp / US_EW, wl = warp % US_EW;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
      if ((int)(g % US_GROUPS) != grp) continue;
      const int mt = tile % n_mt, b = tile / n_mt;
      const int u = g % NBUF;
      umma_tile_epilogue_rl<US_C, 2, US_EW>(d, 0, mt, b, u, (g / NBUF) & 1u, wl, lane, tmem_base, tfull_bar(u), 1, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(u));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- weight-resident kernel, CTA-pair form ---------------------------------------------------------------------------
// The weight-resident kernel above is bound by the tensor core's operand fetch from shared memory (ncu: 70 % of the
// wavefront peak at 38 % tensor pipe).  As a cta_group::2 pair one instruction covers 256 rows (128 per CTA) and each
// CTA supplies only HALF of the B operand: per k step and CTA  x_hi (4 KB) + half of [W_hi | W_lo]  and  x_lo (4 KB) +
// half of W_hi  = 11 -> 9.5 KB at 32 channels, 14 -> 11 KB at 64, with half the MMA instructions per row.  The weights
// stay resident: per tap slot 1 = the CTA's half of the stacked operand (rank 0: the C rows of W_hi, rank 1: the C rows
// of W_lo) and slot 2 = its half of W_hi (C/2 rows) -- 1.5 C rows per tap and CTA, so 64 channels x 11 taps (135 KB) fit
// next to two activation stages, which the single-CTA form (180 KB) cannot hold.  Same MMA order per output element as
// every other kernel of the vocoder's convs: same bits.
template <int US_C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(US_THREADS, 1)
conv1d_umma_wres_pair_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                             const __grid_constant__ CUtensorMap mapBhF, const __grid_constant__ CUtensorMap mapBlF,
                             const __grid_constant__ CUtensorMap mapBhH, const pttspp_conv1d_desc d, const int n_mt,
                             const int n_units, const int rowsA, const int nst) {
  constexpr int NBUF = 512 / (2 * US_C);
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t ROWB = US_C * 2;
  constexpr uint32_t TAPB = (US_C + US_C / 2) * ROWB;  // slot 1 (C rows) + slot 2 (C/2 rows)
  constexpr int MAX_NST = 6;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_bytes = (uint32_t)d.K * TAPB;
  const uint32_t a_plane = (uint32_t)rowsA * ROWB;
  const uint32_t a_stage = ((2u * a_plane) + 1023u) & ~1023u;
  const uint32_t ring = base + ((w_bytes + 1023u) & ~1023u);
  const uint32_t bars_off = ((w_bytes + 1023u) & ~1023u) + (uint32_t)nst * a_stage;
  const uint32_t bars = base + bars_off;
  constexpr int NBARS = 1 + 2 * MAX_NST + 2 * NBUF;
  const uint32_t fullW = bars;
  auto fullA = [&](int st) { return bars + (1 + st) * 8; };
  auto emptyA = [&](int st) { return bars + (1 + MAX_NST + st) * 8; };
  auto tfull_bar = [&](int u) { return bars + (1 + 2 * MAX_NST + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (1 + 2 * MAX_NST + NBUF + u) * 8; };
  const uint32_t tmem_slot = bars + NBARS * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bars_off + NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  constexpr int W_PROD = US_GROUPS * US_EW, W_MMA = W_PROD + 1;
  if (threadIdx.x == 0) {
    mbar_init(fullW, 1);
    for (int st = 0; st < MAX_NST; ++st) {
      mbar_init(fullA(st), 1);
      mbar_init(emptyA(st), 1);
    }
    for (int u = 0; u < NBUF; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), 2 * US_EW);  // one epilogue group of each CTA
    }
    fence_barrier_init();
  }
  if (warp == W_PROD && lane == 0) {
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapBhF);
    tma_prefetch_desc(&mapBlF);
    tma_prefetch_desc(&mapBhH);
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int n_blocks = d.B * n_mt;  // flat (utterance, 128-row block) list; unit u = blocks 2u (rank 0), 2u + 1 (rank 1)

  if (warp == W_PROD) {
    // ================= TMA producer (both CTAs): own weight halves once, own activation block per unit =================
    if (elect_one()) {
      if (rank == 0) mbar_expect_tx(fullW, 2u * w_bytes);
      for (int tap = 0; tap < d.K; ++tap) {
        tma_load_2d_pair(base + (uint32_t)tap * TAPB, rank == 0 ? &mapBhF : &mapBlF, fullW, 0, tap * US_C);
        tma_load_2d_pair(base + (uint32_t)tap * TAPB + US_C * ROWB, &mapBhH, fullW, 0, tap * US_C + (int)rank * (US_C / 2));
      }
    }
    __syncwarp();
    uint32_t g = 0;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters, ++g) {
      const int q = 2 * unit + (int)rank;
      const int b = min(q / n_mt, d.B - 1), mt = (q < n_blocks) ? q - (q / n_mt) * n_mt : n_mt;  // past the end: no rows
      const int st = (int)(g % (uint32_t)nst);
      mbar_wait_warp(emptyA(st), ((g / (uint32_t)nst) & 1u) ^ 1u);
      const int row0 = d.m_begin + mt * UM_BM - d.pad;
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(fullA(st), 4u * a_plane);
        tma_load_3d_pair(ring + (uint32_t)st * a_stage, &mapAh, fullA(st), 0, row0, b);
        tma_load_3d_pair(ring + (uint32_t)st * a_stage + a_plane, &mapAl, fullA(st), 0, row0, b);
      }
      __syncwarp();
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer (leader CTA) =================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * UM_BM, US_C);
      constexpr uint32_t idesc2 = umma_idesc_f16(2 * UM_BM, 2 * US_C);
      const uint64_t descW = (US_C == 32) ? umma_desc_k_sw64(base) : umma_desc_k_sw128(base);
      mbar_wait_warp(fullW, 0);
      tc_fence_after();
      uint32_t g = 0;
      for (int unit = cluster_id; unit < n_units; unit += n_clusters, ++g) {
        const int st = (int)(g % (uint32_t)nst), u = g % NBUF;
        mbar_wait_warp(tempty_bar(u), ((g / NBUF) & 1u) ^ 1u);  // both CTAs' epilogue groups have drained buffer u
        mbar_wait_warp(fullA(st), (g / (uint32_t)nst) & 1u);
        tc_fence_after();
        const uint32_t acc_main = tmem_base + (uint32_t)(u * 2 * US_C);
        const uint32_t acc_cross = acc_main + (uint32_t)US_C;
        const uint64_t descA = (US_C == 32) ? umma_desc_k_sw64(ring + (uint32_t)st * a_stage)
                                            : umma_desc_k_sw128(ring + (uint32_t)st * a_stage);
        if (elect_one()) {
          for (int tap = 0; tap < d.K; ++tap) {
            const uint64_t dAh = descA + (uint64_t)(((uint32_t)(tap * d.dil) * ROWB) >> 4);
            const uint64_t dAl = dAh + (uint64_t)(a_plane >> 4);
            const uint64_t dB1 = descW + (uint64_t)(((uint32_t)tap * TAPB) >> 4);
            const uint64_t dB2 = dB1 + (uint64_t)((US_C * ROWB) >> 4);
#pragma unroll
            for (int kk = 0; kk < US_C / 16; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 2);
              const uint32_t acc = (tap | kk) ? 1u : 0u;
              umma_f16_pair(acc_main, dAh + adv, dB1 + adv, idesc2, acc);  // x_hi . [W_hi | W_lo] -> main | cross
              umma_f16_pair(acc_cross, dAl + adv, dB2 + adv, idesc, 1u);   // x_lo . W_hi
            }
          }
          umma_commit_pair(emptyA(st));
          umma_commit_pair(tfull_bar(u));
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue (both CTAs): two groups of 8 warps alternate units, each CTA drains its own rows ======
    const int grp = warp / US_EW, wl = warp % US_EW;
    uint32_t g = 0;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters, ++g) {
      if ((int)(g % US_GROUPS) != grp) continue;
      const int q = 2 * unit + (int)rank;
      const int b = min(q / n_mt, d.B - 1), mt = (q < n_blocks) ? q - (q / n_mt) * n_mt : n_mt;
      const int u = g % NBUF;
      umma_tile_epilogue_rl<US_C, 2, US_EW>(d, 0, mt, b, u, (g / NBUF) & 1u, wl, lane, tmem_base, tfull_bar(u), 1, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar(u));
        else mbar_arrive_cluster(tempty_bar(u), 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- fused anti-aliased Snake -> weight-resident conv -------------------------------------------------------------
// BigVGAN's AMP layer is  x -> AA-Snake -> conv1 -> AA-Snake -> conv2 (+x)  (vocoders/bigvgan.py:42-47,
// layers/activations.py:22-138).  With the activation as its own launch the activated tensor makes a round trip through
// HBM as split-fp16 operand planes (write 4 B + read 4 B per element, 18 times per stage).  Here the conv kernel's TMA
// producer is replaced by ACTIVATION-PRODUCER warps: they read the fp32 pre-activation rows of the tile (128 rows + the
// conv halo + the activation's own +-5 rows), run the 2x up-sample -> Snake -> 2x down-sample in packed fp32x2 registers
// (aa_math.cuh, the arithmetic of aa_snake_pair_kernel in the same order: results are bit-identical to the two-launch
// path) and write the result straight into the swizzled K-major operand stage that the MMA descriptors read; the
// activated tensor never exists in HBM.  Rows outside [0, L) are the conv's zero padding (written as zeros), the
// activation's own replicate padding is applied on the up-sampled signal exactly as in the stand-alone kernel.
//   warps [0, EW)         epilogue (TMEM -> registers -> bias / residual / MRF average -> HBM)
//   warp  EW              weight TMA (once) + MMA issuer
//   warps (EW, EW + G*NG] NG producer groups of G warps; group j produces tiles j, j + NG, ... of this CTA into the
//                         operand ring (stage = tile sequence number % NST); a thread owns one channel pair and one
//                         contiguous strip of the tile's rows

// one strip of activated rows [t_lo, t_hi) of one channel pair -> operand planes in shared memory
template <int ROWB, int AF_TB>
__device__ __forceinline__ void aa_strip_to_stage(const float* __restrict__ xb, const int ld, const int L, const int t_lo,
                                                  const int t_hi, const int row0, const uint32_t s_hi,
                                                  const uint32_t plane_bytes, const int cp, const f32x2 (&e)[6],
                                                  const f32x2 (&g)[6], const f32x2 a2, const f32x2 inv_alpha) {
  constexpr uint32_t SWZ = (ROWB == 128) ? 7u : 3u;  // SWIZZLE_128B / SWIZZLE_64B: 16-byte chunk ^= address bits 7..
  f32x2 sv[2 * AF_TB + 10], xw[AF_TB + 5], xn[AF_TB];
  {
    // the 10 up-sampled values in front of the strip (indices 2*t_lo - 5 .. 2*t_lo + 4) from x[t_lo - 5 .. t_lo + 4]
    f32x2 xs[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      int l = t_lo - 5 + i;
      l = l < 0 ? 0 : (l > L - 1 ? L - 1 : l);
      xs[i] = *reinterpret_cast<const f32x2*>(xb + (int64_t)l * ld);
    }
#pragma unroll
    for (int i = 0; i < AF_TB; ++i) {
      int l = t_lo + 5 + i;
      l = l > L - 1 ? L - 1 : l;
      xn[i] = *reinterpret_cast<const f32x2*>(xb + (int64_t)l * ld);
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
      if (2 * t_lo - 5 + i < 0) sv[i] = sv[i + 1];
    const int imax = 2 * (L - t_lo) + 4;
#pragma unroll
    for (int i = 1; i < 10; ++i)
      if (i > imax) sv[i] = sv[i - 1];
#pragma unroll
    for (int i = 0; i < 5; ++i) xw[i] = xs[5 + i];
  }
#pragma unroll 1
  for (int tb = t_lo; tb < t_hi; tb += AF_TB) {
#pragma unroll
    for (int i = 0; i < AF_TB; ++i) xw[5 + i] = xn[i];
    if (tb + AF_TB < t_hi) {  // request the next block's inputs now
#pragma unroll
      for (int i = 0; i < AF_TB; ++i) {
        int l = tb + AF_TB + 5 + i;
        l = l > L - 1 ? L - 1 : l;
        xn[i] = *reinterpret_cast<const f32x2*>(xb + (int64_t)l * ld);
      }
    }
    // new up-sampled values: sv[i] <-> index 2*tb - 5 + i, i = 10 .. 2*TB + 9
#pragma unroll
    for (int i = 10; i < 2 * AF_TB + 10; ++i) {
      f32x2 u = pk2(0.f, 0.f);
      if ((i & 1) == 0) {
#pragma unroll
        for (int dd = 0; dd < 6; ++dd) u = fma2(xw[i / 2 + dd - 5], e[dd], u);
      } else {
#pragma unroll
        for (int dd = 0; dd < 6; ++dd) u = fma2(xw[(i - 1) / 2 + dd - 5], e[5 - dd], u);
      }
      sv[i] = snake2(u, a2, inv_alpha);
    }
    if (tb + AF_TB + 4 > L - 1) {
      const int imax = 2 * (L - tb) + 4;  // block index of up-sampled sample 2L-1: later ones repeat it
#pragma unroll
      for (int i = 10; i < 2 * AF_TB + 10; ++i)
        if (i > imax) sv[i] = sv[i - 1];
    }
#pragma unroll
    for (int t = 0; t < AF_TB; ++t) {
      if (tb + t < t_hi) {
        f32x2 acc = pk2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 12; ++k) acc = fma2(sv[2 * t + k], g[k < 6 ? k : 11 - k], acc);
        float a0, a1;
        upk2(acc, a0, a1);
        uint32_t hw, lw;
        split2_f16(a0, a1, hw, lw);
        uint32_t off = (uint32_t)(tb + t - row0) * ROWB + (uint32_t)cp * 4u;
        off ^= ((off >> 7) & SWZ) << 4;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_hi + off), "r"(hw) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_hi + plane_bytes + off), "r"(lw) : "memory");
      }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) sv[i] = sv[2 * AF_TB + i];
#pragma unroll
    for (int i = 0; i < 5; ++i) xw[i] = xw[AF_TB + i];
  }
}

template <int US_C, int EW, int G, int NG, int TB>
__global__ void __launch_bounds__((EW + 1 + G * NG) * 32, 1)
aa_conv_wres_kernel(const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                    const pttspp_conv1d_desc d, const float* __restrict__ log_alpha, const float* __restrict__ up_f,
                    const float* __restrict__ down_f, const int n_mt, const int n_tiles, const int rowsA, const int nst,
                    const int dbg) {
  constexpr int NBUF = 512 / (2 * US_C);  // accumulator buffers (main | cross): all 512 TMEM columns
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t ROWB = US_C * 2;  // bytes per operand row
  constexpr int MAX_NST = 6;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_bytes = 2u * (uint32_t)d.K * US_C * ROWB;  // per tap: C hi rows followed by C lo rows
  const uint32_t a_plane = (uint32_t)rowsA * ROWB;
  const uint32_t a_stage = ((2u * a_plane) + 1023u) & ~1023u;
  const uint32_t ring = base + ((w_bytes + 1023u) & ~1023u);
  const uint32_t bars_off = ((w_bytes + 1023u) & ~1023u) + (uint32_t)nst * a_stage;
  const uint32_t bars = base + bars_off;
  // fullW, fullA[MAX_NST], emptyA[MAX_NST], tfull[NBUF], tempty[NBUF]
  constexpr int NBARS = 1 + 2 * MAX_NST + 2 * NBUF;
  const uint32_t fullW = bars;
  auto fullA = [&](int st) { return bars + (1 + st) * 8; };
  auto emptyA = [&](int st) { return bars + (1 + MAX_NST + st) * 8; };
  auto tfull_bar = [&](int u) { return bars + (1 + 2 * MAX_NST + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (1 + 2 * MAX_NST + NBUF + u) * 8; };
  const uint32_t tmem_slot = bars + NBARS * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bars_off + NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_MMA = EW;
  if (threadIdx.x == 0) {
    mbar_init(fullW, 1);
    for (int st = 0; st < MAX_NST; ++st) {
      mbar_init(fullA(st), G);  // one arrival per producer warp of the group
      mbar_init(emptyA(st), 1);
    }
    for (int u = 0; u < NBUF; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), EW);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    if (lane == 0) {
      tma_prefetch_desc(&mapBh);
      tma_prefetch_desc(&mapBl);
    }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == W_MMA) {
    // ================= weights (once) + MMA issuer =================
    if (elect_one()) {
      mbar_expect_tx(fullW, w_bytes);
      for (int tap = 0; tap < d.K; ++tap) {
        tma_load_2d(base + (uint32_t)tap * (2 * US_C * ROWB), &mapBh, fullW, 0, tap * US_C);
        tma_load_2d(base + (uint32_t)tap * (2 * US_C * ROWB) + US_C * ROWB, &mapBl, fullW, 0, tap * US_C);
      }
    }
    __syncwarp();
    constexpr uint32_t idesc = umma_idesc_f16(UM_BM, US_C);
    constexpr uint32_t idesc2 = umma_idesc_f16(UM_BM, 2 * US_C);
    const uint64_t descW = (US_C == 32) ? umma_desc_k_sw64(base) : umma_desc_k_sw128(base);
    mbar_wait_warp(fullW, 0);
    tc_fence_after();
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
      const int st = (int)(g % (uint32_t)nst), u = g % NBUF;
      mbar_wait_warp(tempty_bar(u), ((g / NBUF) & 1u) ^ 1u);
      mbar_wait_warp_sleep(fullA(st), (g / (uint32_t)nst) & 1u, 64);
      tc_fence_after();
      const uint32_t acc_main = tmem_base + (uint32_t)(u * 2 * US_C);
      const uint32_t acc_cross = acc_main + (uint32_t)US_C;
      const uint64_t descA = (US_C == 32) ? umma_desc_k_sw64(ring + (uint32_t)st * a_stage)
                                          : umma_desc_k_sw128(ring + (uint32_t)st * a_stage);
      if (elect_one()) {
        for (int tap = 0; tap < d.K; ++tap) {
          const uint64_t dAh = descA + (uint64_t)(((uint32_t)(tap * d.dil) * ROWB) >> 4);  // taps share the halo block
          const uint64_t dAl = dAh + (uint64_t)(a_plane >> 4);
          const uint64_t dBh = descW + (uint64_t)(((uint32_t)tap * (2 * US_C * ROWB)) >> 4);
#pragma unroll
          for (int kk = 0; kk < US_C / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);
            const uint32_t acc = (tap | kk) ? 1u : 0u;
            umma_f16(acc_main, dAh + adv, dBh + adv, idesc2, acc);  // x_hi . [W_hi | W_lo] -> main | cross
            umma_f16(acc_cross, dAl + adv, dBh + adv, idesc, 1u);   // x_lo . W_hi
          }
        }
        umma_commit(emptyA(st));
        umma_commit(tfull_bar(u));
      }
      __syncwarp();
    }
  } else if (warp < EW) {
    // ================= epilogue =================
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
      const int mt = tile % n_mt, b = tile / n_mt;
      const int u = g % NBUF;
      // the producers are issue bound and this warp has NBUF accumulator buffers of slack: request the tile's residual
      // rows into L2, then sleep-poll (the epilogue's own wait below passes at once)
      epi_rl_prefetch_tile<US_C, EW>(d, 0, mt, b, warp, lane);
      mbar_wait_warp_sleep(tfull_bar(u), (g / NBUF) & 1u, 256);
      umma_tile_epilogue_rl<US_C, 2, EW>(d, 0, mt, b, u, (g / NBUF) & 1u, warp, lane, tmem_base, tfull_bar(u), 1,
                                         (dbg & 2) ? 64 : 0);  // timing experiment: no epilogue work
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(u));
    }
  } else {
    // ================= activation producers =================
    const int aw = warp - (EW + 1);
    const int grp = aw / G;
    const int tid_g = (aw % G) * 32 + lane;
    constexpr int P = US_C / 2;  // channel pairs
    constexpr int NSTRIPS = G * 32 / P;
    const int cp = tid_g % P, strip = tid_g / P;
    const int L = d.T_in;
    const int RN = UM_BM + (d.K - 1) * d.dil;  // operand rows the taps read
    const int slen = (RN + NSTRIPS - 1) / NSTRIPS;
    f32x2 e[6], gd[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const float ef = 2.f * up_f[10 - 2 * q];
      e[q] = pk2(ef, ef);
      gd[q] = pk2(down_f[q], down_f[q]);
    }
    const float al0 = expf(log_alpha[2 * cp]), al1 = expf(log_alpha[2 * cp + 1]);
    const f32x2 inv_alpha = pk2(1.f / (al0 + 1e-9f), 1.f / (al1 + 1e-9f));
    const f32x2 a2 = pk2(al0, al1);
    for (uint32_t g = (uint32_t)grp;; g += NG) {
      const long long tile_ll = (long long)blockIdx.x + (long long)g * gridDim.x;
      if (tile_ll >= n_tiles) break;
      const int tile = (int)tile_ll;
      const int st = (int)(g % (uint32_t)nst);
      mbar_wait_warp(emptyA(st), ((g / (uint32_t)nst) & 1u) ^ 1u);
      const int mt = tile % n_mt, b = tile / n_mt;
      const int row0 = d.m_begin + mt * UM_BM - d.pad;  // utterance row of operand row 0
      const int t_begin = row0 + strip * slen;
      const int t_end = min(row0 + RN, t_begin + slen);
      const int t_lo = max(t_begin, 0), t_hi = min(t_end, L);
      const uint32_t s_hi = ring + (uint32_t)st * a_stage;
      // rows outside the utterance: the conv's zero padding
      for (int t = t_begin; t < t_end; ++t) {
        if (t >= t_lo && t < t_hi) continue;
        uint32_t off = (uint32_t)(t - row0) * ROWB + (uint32_t)cp * 4u;
        off ^= ((off >> 7) & ((ROWB == 128) ? 7u : 3u)) << 4;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_hi + off), "r"(0u) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_hi + a_plane + off), "r"(0u) : "memory");
      }
      if (t_lo < t_hi && !(dbg & 1))  // (dbg & 1: timing experiment without the activation)
        aa_strip_to_stage<(int)ROWB, TB>(d.in + (int64_t)b * d.in_bs + 2 * cp, d.in_ld, L, t_lo, t_hi, row0, s_hi, a_plane, cp,
                                     e, gd, a2, inv_alpha);
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy operand reads
      __syncwarp();
      if (lane == 0) mbar_arrive(fullA(st));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- probe: UMMA shared-memory descriptors whose start row is not a multiple of 8 --------------------------------
// D[128][128] = A[row_off : row_off + 128][0:64] . B[0:128][0:64]^T with the A tile loaded ONCE (144 rows) and the
// descriptor start advanced by row_off * 128 bytes.  mode 0: base_offset field 0; mode 1: base_offset =
// (start_address >> 7) & 7 as the PTX ISA prescribes for starts that are not aligned to the 1024-byte swizzle
// pattern.  Decides whether a conv can share one halo tile between its taps.
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int row_off,
                  int mode, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 144 * 128;  // 18432 = 18 * 1024: B stays 1024-aligned
  const uint32_t bar_full = base + 144 * 128 + 128 * 128, bar_done = bar_full + 8, slot = bar_full + 16;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + 144 * 128 + 128 * 128 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_full, 144 * 128 + 128 * 128);
    tma_load_2d(sA, &mapA, bar_full, 0, 0);
    tma_load_2d(sB, &mapB, bar_full, 0, 0);
    mbar_wait(bar_full, 0);
    tc_fence_after();
    const uint32_t a_start = sA + (uint32_t)row_off * 128u;
    uint64_t dA = umma_desc_k_sw128(a_start);
    if (mode == 1) dA |= (uint64_t)((a_start >> 7) & 7u) << 49;
    const uint64_t dB = umma_desc_k_sw128(sB);
    for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, dA + (uint64_t)(kk * 2), dB + (uint64_t)(kk * 2), umma_idesc_f16(128, 128), kk != 0);
    umma_commit(bar_done);
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  for (int c = 0; c < 128; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * 128 + c + e] = v[e];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Developer knobs (kernel-variant selection for the op tests and A/B measurements) are read from the environment ONCE
// and cached; pttspp_debug_reload_env() re-reads them (tests that switch variants inside one process call it).
struct UmmaEnv {
  char pair = 0, rl = 0, nacc = 0, epi = 0;  // first character of the variable, 0 when unset
  bool no_tma_store = false, debug = false, no_wres64 = false, a_stationary = false;
  int order = 0;
  bool longk_narrow = false;
  int chunk_iters = 4;  // chunked mode: (slab, tap) iterations per accumulator buffer (4 x 64 channels = 16 MMA steps)
  void load() {
    auto first = [](const char* name) -> char { const char* e = getenv(name); return e ? e[0] : (char)0; };
    pair = first("PTTSPP_UMMA_PAIR");
    rl = first("PTTSPP_UMMA_RL");
    nacc = first("PTTSPP_UMMA_NACC");
    epi = first("PTTSPP_UMMA_EPI");
    no_tma_store = getenv("PTTSPP_UMMA_NO_TMA_STORE") != nullptr;
    debug = getenv("PTTSPP_UMMA_DEBUG") != nullptr;
    no_wres64 = getenv("PTTSPP_UMMA_NO_WRES64") != nullptr;
    a_stationary = getenv("PTTSPP_UMMA_AS") != nullptr;
    const char* oe = getenv("PTTSPP_UMMA_ORDER");  // experiments: bits 0-1 MMA order, 4 no operand loads, 8 no stores, 16 no L2 prefetch
    order = oe ? atoi(oe) : 0;
    const char* lk = getenv("PTTSPP_UMMA_LONGK");
    longk_narrow = lk && lk[0] == 'n';
    const char* ce = getenv("PTTSPP_UMMA_CHUNK");
    chunk_iters = ce ? std::max(1, atoi(ce)) : 4;
  }
};
std::mutex g_env_mutex;
UmmaEnv& umma_env_storage() {
  static UmmaEnv e = [] { UmmaEnv x; x.load(); return x; }();
  return e;
}
UmmaEnv umma_env() {
  std::lock_guard<std::mutex> lk(g_env_mutex);
  return umma_env_storage();
}

// cuTensorMapEncodeTiled costs microseconds and the same (pointer, geometry) recurs on every step of the sampling loop
// and every call with a stable workspace: encoded maps are cached (bounded; cleared when full).
struct MapKey {
  const void* ptr;
  uint64_t dims[3], strides[2];
  uint32_t box[3];
  int rank, dtype, swz;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
static_assert(sizeof(MapKey) % 8 == 0, "MapKey is hashed as 64-bit words");

CUtensorMap make_map(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                     CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                     CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rank = rank; key.dtype = (int)dtype; key.swz = (int)swz;
  for (int i = 0; i < rank; ++i) key.dims[i] = dims[i], key.box[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  EncodeTiledFn fn = encode_fn();
  PT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap m;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t gbox[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i], gbox[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(&m, dtype, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rank %d, dims %llu x %llu)", (int)r, rank,
           (unsigned long long)dims[0], (unsigned long long)dims[1]);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() >= 8192) cache.clear();
    cache.emplace(key, m);
  }
  return m;
}

// Per-device launch state: the multiprocessor count, and which kernels have opted in to 227 KB of dynamic shared memory
// on that device (cudaFuncSetAttribute is per device; a process may drive several GPUs from several threads).
constexpr int kMaxDevices = 64;
struct DeviceState {
  int num_sms = 0;
  std::unordered_map<const void*, int> kernel_setup;  // kernel -> resident clusters (pair kernels) or 1
};
std::mutex g_dev_mutex;
DeviceState g_dev[kMaxDevices];
int current_device() {
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  PT_CHECK(dev >= 0 && dev < kMaxDevices, "device ordinal %d out of range", dev);
  return dev;
}
int device_num_sms() {
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  if (g_dev[dev].num_sms == 0) PT_CUDA(cudaDeviceGetAttribute(&g_dev[dev].num_sms, cudaDevAttrMultiProcessorCount, dev));
  return g_dev[dev].num_sms;
}
// opt the kernel in to `smem` bytes of dynamic shared memory once per device
void ensure_smem_optin(const void* kern, int smem) {
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  auto& m = g_dev[dev].kernel_setup;
  if (m.find(kern) != m.end()) return;
  PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  m.emplace(kern, 1);
}

constexpr int UM_BN = 128;
constexpr int UM_STAGES = 3;

// TMA bulk stores need the plain row mapping (row = m) over the whole output and 16-byte aligned strides
bool tma_out_ok(const pttspp_conv1d_desc& d) {
  if (!conv_epilogue_vec_ok(d)) return false;
  if (d.out_mul != 1 || d.out_off != 0 || d.m_begin != 0 || d.M != d.T_out) return false;
  if (d.out && (d.out_ld % 4 != 0 || d.out_bs % 4 != 0)) return false;
  if (d.out_hi && (d.out_plane_ld % 8 != 0 || d.out_plane_bs % 8 != 0)) return false;
  return !umma_env().no_tma_store;
}

OutMaps make_out_maps(const pttspp_conv1d_desc& d) {
  OutMaps om;
  memset(&om, 0, sizeof(om));
  const bool gate = d.act == PTTSPP_ACT_GATE;
  const uint64_t cols = gate ? d.Cout / 2 : d.Cout;
  const uint32_t box[3] = {gate ? 8u : 16u, 32u, 1u};
  const uint64_t dims[3] = {cols, (uint64_t)d.T_out, (uint64_t)d.B};
  if (d.out) {
    const uint64_t str[2] = {(uint64_t)d.out_ld * 4, (uint64_t)d.out_bs * 4};
    om.f32 = make_map(d.out, 3, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE);
  }
  if (d.out_hi) {
    const uint64_t str[2] = {(uint64_t)d.out_plane_ld * 2, (uint64_t)d.out_plane_bs * 2};
    om.hi = make_map(d.out_hi, 3, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_NONE);
    om.lo = make_map(d.out_lo, 3, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_SWIZZLE_NONE);
  }
  return om;
}


// coalescing epilogue preconditions: the vector-path alignment rules; the gate activation only without residual / beta
bool epilogue_co_ok(const pttspp_conv1d_desc& d) {
  if (!conv_epilogue_vec_ok(d)) return false;
  if (d.act == PTTSPP_ACT_GATE && (d.res || (d.out && d.beta != 0.f))) return false;
  return true;
}

// row-per-lane 256-bit epilogue: 32-byte aligned bases / row strides / batch strides
bool epilogue_rl_ok(const pttspp_conv1d_desc& d) {
  if (!epilogue_co_ok(d)) return false;
  auto a32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
  auto okf = [&](const void* p, int64_t bs, int ld) { return !p || (a32(p) && bs % 8 == 0 && ld % 8 == 0); };
  if (!okf(d.addend, d.addend_bs, d.addend_ld) || !okf(d.res, d.res_bs, d.res_ld) || !okf(d.out, d.out_bs, d.out_ld))
    return false;
  if (d.out_hi && !(a32(d.out_hi) && a32(d.out_lo) && d.out_plane_bs % 16 == 0 && d.out_plane_ld % 16 == 0)) return false;
  if (d.res_hi && !(d.res_lo && !d.res && a32(d.res_hi) && a32(d.res_lo) && d.res_plane_bs % 16 == 0 &&
                    d.res_plane_ld % 16 == 0 && (!d.res_plane_sub || aligned16(d.res_plane_sub))))
    return false;
  return true;
}

// A 1x1 contraction over densely packed [B][T] rows is one [B*T]-row problem: no halo, no per-utterance tile tails.
bool flatten_batch_ok(const pttspp_conv1d_desc& d) {
  if (d.K != 1 || d.pad != 0 || d.m_begin != 0 || d.M != d.T_out || d.M != d.T_in || d.out_mul != 1 || d.out_off != 0)
    return false;
  if (d.out_len || d.in_len) return false;
  if (d.in_bs != (int64_t)d.T_in * d.in_ld) return false;
  if (d.out && d.out_bs != (int64_t)d.T_out * d.out_ld) return false;
  if (d.res && d.res_bs != (int64_t)d.T_out * d.res_ld) return false;
  if (d.addend && d.addend_bs != (int64_t)d.T_out * d.addend_ld) return false;
  if (d.out_hi && d.out_plane_bs != (int64_t)d.T_out * d.out_plane_ld) return false;
  if (d.res_hi && d.res_plane_bs != (int64_t)d.T_out * d.res_plane_ld) return false;
  return (int64_t)d.B * d.T_in < (1ll << 30);
}
void flatten_batch(pttspp_conv1d_desc& d) {
  const int rows = d.B * d.T_in;
  d.T_in = d.T_out = d.M = rows;
  d.B = 1;
}

int pair_clusters(const void* kern, size_t smem, int threads) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

// one-time setup per kernel instantiation: opt in to 227 KB of dynamic shared memory, query resident clusters
template <int NBST, int EW, int NACC>
int pair_kernel_setup(int num_sms, bool debug) {
  auto kern = conv1d_umma_pair_kernel<NBST, EW, NACC>;
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  auto& m = g_dev[dev].kernel_setup;
  auto it = m.find((const void*)kern);
  if (it != m.end()) return it->second;
  PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int n = pair_clusters((const void*)kern, 227 * 1024, (EW + 2) * 32);
  if (debug) fprintf(stderr, "[pttspp] pair kernel<%d>: cudaOccupancyMaxActiveClusters -> %d\n", NBST, n);
  const int max_clusters = (n > 0) ? std::min(n, num_sms / 2) : num_sms / 2;
  m.emplace((const void*)kern, max_clusters);
  return max_clusters;
}

// Pair kernel launch; returns false when the shape does not qualify (caller falls back to the streaming kernel).
bool conv1d_umma_pair_launch(pttspp_conv1d_desc d, pttspp_conv1d_desc d2, bool dual, int cout1, int total_cout,
                             cudaStream_t s, int num_sms) {
  const UmmaEnv ev = umma_env();
  const char env_pair = ev.pair;
  const bool needs_pair = d.res_hi != nullptr || (dual && d2.res_hi != nullptr);  // only this kernel's epilogue reads it
  if (env_pair == '0' && !needs_pair) return false;
  if (!epilogue_co_ok(d) || (dual && !epilogue_co_ok(d2))) return false;
  if (total_cout % UP_BN != 0 || d.Cin % UM_BK != 0 || d.Cin / UM_BK > UP_MAX_SLAB) return false;
  // impl 4 (the vocoder: 1e-4 RMS waveform bar, measured 8e-6): contractions of up to 256 accumulations per accumulator
  // and single-N-tile convs take this kernel too -- it loads the activation block once per 256-row unit and each CTA
  // only half of every weight tile, where the streaming kernels are bound by the L2 -> SM fill (ncu: 5.4-10 TB/s of
  // re-streamed weight / activation tiles); everything else keeps the chunked streaming kernel for long contractions
  const bool longk = d.impl == 4;
  if (d.K * d.Cin / 16 > (longk ? 256 : 64)) return false;
  if (flatten_batch_ok(d) && (!dual || flatten_batch_ok(d2))) {
    flatten_batch(d);
    if (dual) flatten_batch(d2);
  }
  const int nslab = d.Cin / UM_BK;
  const int rowsA = UM_BM + round_up((d.K - 1) * d.dil, 8);
  if (rowsA > 256) return false;
  const size_t a_bytes = (size_t)nslab * 2 * rowsA * 128;
  const int ew = 16;  // epilogue warps (8 measured slower: tools/bench_conv.py history in profiles/)
  // row-per-lane 256-bit epilogue (no staging tile) whenever the alignment allows it; PTTSPP_UMMA_RL=0: coalescing one
  const bool epi_rl = epilogue_rl_ok(d) && (!dual || epilogue_rl_ok(d2)) && (needs_pair || ev.rl != '0');
  PT_CHECK(!needs_pair || epi_rl, "conv1d: a residual from operand planes needs 32-byte aligned tensors (row-per-lane epilogue)");
  const size_t fixed = a_bytes + (epi_rl ? 0 : (size_t)ew * 2048) + (2 * UP_MAX_SLAB + 2 * 6 + 8) * 8 + 16 + 1024;
  const size_t cap = 227 * 1024;
  if (fixed + 3 * UP_BST_BYTES > cap) return false;
  const int nbst = (int)std::min<size_t>(6, (cap - fixed) / UP_BST_BYTES);
  const int n_mt = ceil_div(d.M, UM_BM), n_nt = total_cout / UP_BN;  // 128-row blocks per utterance
  const long long n_units = ceil_div64((long long)n_mt * d.B, 2);      // pairs of blocks
  const long long n_tiles = n_units * n_nt;
  if (n_tiles >= (1ll << 30)) return false;
  // worth it only when every pair gets at least a couple of tiles (the activation block is loaded per unit)
  // and when the activation block is reused by at least two N tiles (otherwise the streaming kernel is faster)
  if (env_pair != '2' && !needs_pair && (n_tiles < num_sms || (n_nt < 2 && !(longk && d.K >= 3)))) return false;

  const uint64_t wdims[2] = {(uint64_t)d.Cin, (uint64_t)d.K * total_cout};
  const uint64_t wstr[1] = {(uint64_t)d.Cin * 2};
  const uint32_t wbox[2] = {UM_BK, UP_BHALF};
  const CUtensorMap mBh = make_map(d.w_hi, 2, wdims, wstr, wbox);
  const CUtensorMap mBl = make_map(d.w_lo, 2, wdims, wstr, wbox);
  const uint64_t adims[3] = {(uint64_t)d.Cin, (uint64_t)d.T_in, (uint64_t)d.B};
  const uint64_t astr[2] = {(uint64_t)d.in_ld * 2, (uint64_t)d.in_bs * 2};
  const uint32_t abox[3] = {UM_BK, (uint32_t)rowsA, 1};
  const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox);
  const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox);
  const size_t smem = fixed + (size_t)nbst * UP_BST_BYTES;

  const bool debug = ev.debug;
  auto launch = [&](auto kern, int max_clusters) {
    const int n_clusters = (int)std::min<long long>(n_tiles, std::min(max_clusters, num_sms / 2));
    int cout_first = dual ? cout1 : total_cout, cout_all = total_cout, a_n_mt2 = n_mt, a_n_nt = n_nt, a_tiles = (int)n_tiles,
        a_rowsA = rowsA;
    int a_order = ev.order, a_rl = epi_rl ? 1 : 0;
    void* args[] = {(void*)&mAh, (void*)&mAl, (void*)&mBh, (void*)&mBl, (void*)&d, (void*)&d2, &cout_first, &cout_all,
                    &a_n_mt2, &a_n_nt, &a_tiles, &a_rowsA, &a_order, &a_rl};
    if (debug)
      fprintf(stderr, "[pttspp] pair launch: clusters %d tiles %lld (units %d x nt %d) rowsA %d nbst %d smem %zu rl %d\n",
              n_clusters, n_tiles, (int)n_units, n_nt, rowsA, nbst, smem, (int)epi_rl);
    cudaGetLastError();  // a stale error of an earlier call must not be attributed to this launch
    PT_CUDA(cudaLaunchKernel((const void*)kern, dim3(2 * n_clusters), dim3((ew + 2) * 32), args, smem, s));
    ++g_launch_count;
  };
  // short contractions: one accumulator per tile, four tile buffers (see the kernel)
  const char nae = ev.nacc;
  // K*Cin <= 768 (144 accumulations): measured at cfg2 scale, the one-accumulator mode changes the mel by 1.9e-5 max-abs
  // (bar 1e-3) and takes 4.6 % off the step (tools/nacc_experiment.py); PTTSPP_UMMA_NACC=2 forces two accumulators
  // Only the gated (DiffNet) conv takes the extended range: BigVGAN's plain convs must stay bit-identical between the
  // pair and the streaming kernel (an utterance synthesised alone equals the same utterance inside a batch).
  const int one_acc_limit = (d.act == PTTSPP_ACT_GATE) ? 768 : 256;
  const bool one_acc = ((d.K * d.Cin <= one_acc_limit) && nae != '2') || nae == '1';
#define PT_PAIR_CASE(N, E, A) launch(conv1d_umma_pair_kernel<N, E, A>, pair_kernel_setup<N, E, A>(num_sms, debug))
#define PT_PAIR_SWITCH(E, A)               \
  switch (nbst) {                          \
    case 3: PT_PAIR_CASE(3, E, A); break;  \
    case 4: PT_PAIR_CASE(4, E, A); break;  \
    case 5: PT_PAIR_CASE(5, E, A); break;  \
    default: PT_PAIR_CASE(6, E, A); break; \
  }
  if (one_acc) { PT_PAIR_SWITCH(16, 1) } else { PT_PAIR_SWITCH(16, 2) }
#undef PT_PAIR_SWITCH
#undef PT_PAIR_CASE
  return true;
}


// weight-resident kernel: 32 channels with up to 11 taps, 64 channels with up to 7 (weights + >= 2 halo stages fit)
bool conv1d_umma_c32_ok(const pttspp_conv1d_desc& d) {
  if (!((d.Cin == 32 && d.Cout == 32 && d.K <= US_MAXK) || (d.Cin == 64 && d.Cout == 64 && d.K <= 7))) return false;
  if (d.Cin == 64 && umma_env().no_wres64) return false;
  return d.K >= 1 && d.in_hi && d.in_lo && d.w_hi && d.w_lo && d.in_stride == 1 && !d.in_len && !d.in_add &&
         d.in_ld % 8 == 0 && d.in_bs % 8 == 0 && aligned16(d.in_hi) && aligned16(d.in_lo) && aligned16(d.w_hi) &&
         aligned16(d.w_lo) && d.w_scale_inv > 0.f && UM_BM + round_up((d.K - 1) * d.dil, 8) <= 256 && epilogue_rl_ok(d);
}

template <int C, int NST>
void conv1d_umma_wres_launch_t(const pttspp_conv1d_desc& d, const CUtensorMap& mAh, const CUtensorMap& mAl,
                               const CUtensorMap& mBh, const CUtensorMap& mBl, int rowsA, size_t smem, int num_sms,
                               cudaStream_t s) {
  auto kern = conv1d_umma_c32_kernel<C, NST>;
  ensure_smem_optin((const void*)kern, 227 * 1024);
  const int n_mt = ceil_div(d.M, UM_BM);
  const long long n_tiles = (long long)n_mt * d.B;
  PT_CHECK(n_tiles < (1ll << 30), "conv1d: too many tiles");
  const int grid = (int)std::min<long long>(n_tiles, num_sms);
  kern<<<grid, US_THREADS, smem, s>>>(mAh, mAl, mBh, mBl, d, n_mt, (int)n_tiles, rowsA);
  PT_LAUNCHED();
}

void conv1d_umma_c32_launch(const pttspp_conv1d_desc& d_in, cudaStream_t s) {
  pttspp_conv1d_desc d = d_in;
  d.acc_scale = d_in.acc_scale * d_in.w_scale_inv;
  const int num_sms = device_num_sms();
  const int C = d.Cin;
  const int rowb = C * 2;
  const CUtensorMapSwizzle swz = (C == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  const int rowsA = UM_BM + round_up((d.K - 1) * d.dil, 8);
  const uint64_t wdims[2] = {(uint64_t)C, (uint64_t)d.K * C};
  const uint64_t wstr[1] = {(uint64_t)C * 2};
  const uint32_t wbox[2] = {(uint32_t)C, (uint32_t)C};
  const CUtensorMap mBh = make_map(d.w_hi, 2, wdims, wstr, wbox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const CUtensorMap mBl = make_map(d.w_lo, 2, wdims, wstr, wbox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const uint64_t adims[3] = {(uint64_t)C, (uint64_t)d.T_in, (uint64_t)d.B};
  const uint64_t astr[2] = {(uint64_t)d.in_ld * 2, (uint64_t)d.in_bs * 2};
  const uint32_t abox[3] = {(uint32_t)C, (uint32_t)rowsA, 1};
  const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const size_t w_bytes = round_up(2 * d.K * C * rowb, 1024);
  const size_t a_stage = round_up(2 * rowsA * rowb, 1024);
  const size_t misc = (1 + 2 * 6 + 2 * US_MAX_NBUF) * 8 + 16 + 1024;
  const size_t cap = 227 * 1024;
  PT_CHECK(w_bytes + 2 * a_stage + misc <= cap, "conv1d weight-resident kernel: shared memory budget exceeded");
  int nst = (int)std::min<size_t>(6, (cap - w_bytes - misc) / a_stage);
  if (nst == 5) nst = 4;  // instantiated ring depths: 2, 3, 4, 6
  const size_t smem = w_bytes + (size_t)nst * a_stage + misc;
#define PT_WRES(CC, N) conv1d_umma_wres_launch_t<CC, N>(d, mAh, mAl, mBh, mBl, rowsA, smem, num_sms, s)
  if (C == 32) {
    PT_WRES(32, 6);  // always fits: 44 KB of weights + 6 x 23.5 KB
  } else {
    switch (nst) {
      case 2: PT_WRES(64, 2); break;
      case 3: PT_WRES(64, 3); break;
      case 4: PT_WRES(64, 4); break;
      default: PT_WRES(64, 6); break;
    }
  }
#undef PT_WRES
}

// CTA-pair form of the weight-resident kernel: impl 4 convs (the vocoder) with enough 128-row blocks to fill the machine
bool conv1d_umma_wres_pair_try(const pttspp_conv1d_desc& d_in, cudaStream_t s) {
  static const bool off = getenv("PTTSPP_UMMA_NO_WRES_PAIR") != nullptr;
  const pttspp_conv1d_desc& q = d_in;
  if (off || q.impl != 4) return false;
  if (!((q.Cin == 32 && q.Cout == 32 && q.K <= US_MAXK) || (q.Cin == 64 && q.Cout == 64 && q.K <= US_MAXK))) return false;
  if (!(q.K >= 1 && q.in_hi && q.in_lo && q.w_hi && q.w_lo && q.in_stride == 1 && !q.in_len && !q.in_add &&
        q.in_ld % 8 == 0 && q.in_bs % 8 == 0 && aligned16(q.in_hi) && aligned16(q.in_lo) && aligned16(q.w_hi) &&
        aligned16(q.w_lo) && q.w_scale_inv > 0.f && epilogue_rl_ok(q)))
    return false;
  const int C = q.Cin, rowb = C * 2;
  const int rowsA = UM_BM + round_up((q.K - 1) * q.dil, 8);
  if (rowsA > 256) return false;
  const size_t w_bytes = round_up(q.K * (C + C / 2) * rowb, 1024);
  const size_t a_stage = round_up(2 * rowsA * rowb, 1024);
  const size_t misc = (1 + 2 * 6 + 2 * US_MAX_NBUF) * 8 + 16 + 1024;
  const size_t cap = 227 * 1024;
  if (w_bytes + 2 * a_stage + misc > cap) return false;
  const int num_sms = device_num_sms();
  const int n_mt = ceil_div(q.M, UM_BM);
  const long long n_blocks = (long long)n_mt * q.B;
  if (n_blocks < 2ll * num_sms || n_blocks >= (1ll << 30)) return false;  // every pair gets >= 2 units
  pttspp_conv1d_desc d = d_in;
  d.acc_scale = d_in.acc_scale * d_in.w_scale_inv;
  const CUtensorMapSwizzle swz = (C == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  const uint64_t wdims[2] = {(uint64_t)C, (uint64_t)d.K * C};
  const uint64_t wstr[1] = {(uint64_t)C * 2};
  const uint32_t wboxF[2] = {(uint32_t)C, (uint32_t)C}, wboxH[2] = {(uint32_t)C, (uint32_t)(C / 2)};
  const CUtensorMap mBhF = make_map(d.w_hi, 2, wdims, wstr, wboxF, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const CUtensorMap mBlF = make_map(d.w_lo, 2, wdims, wstr, wboxF, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const CUtensorMap mBhH = make_map(d.w_hi, 2, wdims, wstr, wboxH, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const uint64_t adims[3] = {(uint64_t)C, (uint64_t)d.T_in, (uint64_t)d.B};
  const uint64_t astr[2] = {(uint64_t)d.in_ld * 2, (uint64_t)d.in_bs * 2};
  const uint32_t abox[3] = {(uint32_t)C, (uint32_t)rowsA, 1};
  const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  int nst = (int)std::min<size_t>(6, (cap - w_bytes - misc) / a_stage);
  const size_t smem = w_bytes + (size_t)nst * a_stage + misc;
  const int n_units = (int)ceil_div64(n_blocks, 2);
  const int n_clusters = std::min(n_units, num_sms / 2);
  int a_n_mt = n_mt, a_units = n_units, a_rowsA = rowsA, a_nst = nst;
  void* args[] = {(void*)&mAh, (void*)&mAl, (void*)&mBhF, (void*)&mBlF, (void*)&mBhH, (void*)&d, &a_n_mt, &a_units, &a_rowsA,
                  &a_nst};
  const void* kern = (C == 32) ? (const void*)conv1d_umma_wres_pair_kernel<32> : (const void*)conv1d_umma_wres_pair_kernel<64>;
  ensure_smem_optin(kern, 227 * 1024);
  cudaGetLastError();
  PT_CUDA(cudaLaunchKernel(kern, dim3(2 * n_clusters), dim3(US_THREADS), args, smem, s));
  ++g_launch_count;
  return true;
}

// fused AA-Snake -> conv: the weight-resident geometries, fp32 pre-activation input
bool aa_conv1d_ok(const pttspp_conv1d_desc& d) {
  if (!((d.Cin == 32 && d.Cout == 32 && d.K <= US_MAXK) || (d.Cin == 64 && d.Cout == 64 && d.K <= 7))) return false;
  const int C = d.Cin, rowb = C * 2;
  const int rowsA = UM_BM + round_up((d.K - 1) * d.dil, 8);
  const size_t w_bytes = round_up(2 * d.K * C * rowb, 1024), a_stage = round_up(2 * rowsA * rowb, 1024);
  if (w_bytes + 2 * a_stage + 4096 > 227 * 1024) return false;
  return d.K >= 1 && d.in && (reinterpret_cast<uintptr_t>(d.in) & 7) == 0 && d.in_ld % 2 == 0 && d.in_bs % 2 == 0 &&
         d.w_hi && d.w_lo && d.in_stride == 1 && !d.in_len && !d.in_add && aligned16(d.w_hi) && aligned16(d.w_lo) &&
         d.w_scale_inv > 0.f && rowsA <= 256 && d.act != PTTSPP_ACT_GATE && epilogue_rl_ok(d);
}

template <int C, int EW, int G, int NG, int TB>
void aa_conv1d_launch_t(const pttspp_conv1d_desc& d, const CUtensorMap& mBh, const CUtensorMap& mBl, const float* log_alpha,
                        const float* up_f, const float* down_f, int rowsA, int nst, size_t smem, int num_sms,
                        cudaStream_t s) {
  auto kern = aa_conv_wres_kernel<C, EW, G, NG, TB>;
  ensure_smem_optin((const void*)kern, 227 * 1024);
  const int n_mt = ceil_div(d.M, UM_BM);
  const long long n_tiles = (long long)n_mt * d.B;
  PT_CHECK(n_tiles < (1ll << 30), "conv1d: too many tiles");
  const int grid = (int)std::min<long long>(n_tiles, num_sms);
  static const int dbg = [] { const char* e = getenv("PTTSPP_AAF_DBG"); return e ? atoi(e) : 0; }();  // timing experiments
  kern<<<grid, (EW + 1 + G * NG) * 32, smem, s>>>(mBh, mBl, d, log_alpha, up_f, down_f, n_mt, (int)n_tiles, rowsA, nst, dbg);
  PT_LAUNCHED();
}

void aa_conv1d_launch(const pttspp_conv1d_desc& d_in, const float* log_alpha, const float* up_f, const float* down_f,
                      cudaStream_t s) {
  pttspp_conv1d_desc d = d_in;
  d.acc_scale = d_in.acc_scale * d_in.w_scale_inv;
  const int num_sms = device_num_sms();
  const int C = d.Cin;
  const int rowb = C * 2;
  const CUtensorMapSwizzle swz = (C == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  const int rowsA = UM_BM + round_up((d.K - 1) * d.dil, 8);
  const uint64_t wdims[2] = {(uint64_t)C, (uint64_t)d.K * C};
  const uint64_t wstr[1] = {(uint64_t)C * 2};
  const uint32_t wbox[2] = {(uint32_t)C, (uint32_t)C};
  const CUtensorMap mBh = make_map(d.w_hi, 2, wdims, wstr, wbox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const CUtensorMap mBl = make_map(d.w_lo, 2, wdims, wstr, wbox, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, swz);
  const size_t w_bytes = round_up(2 * d.K * C * rowb, 1024);
  const size_t a_stage = round_up(2 * rowsA * rowb, 1024);
  const size_t misc = (1 + 2 * 6 + 2 * US_MAX_NBUF) * 8 + 16 + 1024;
  const size_t cap = 227 * 1024;
  PT_CHECK(w_bytes + 2 * a_stage + misc <= cap, "aa_conv1d: shared memory budget exceeded");
  const int nst = (int)std::min<size_t>(4, (cap - w_bytes - misc) / a_stage);
  const size_t smem = w_bytes + (size_t)nst * a_stage + misc;
  static const int cfg = [] { const char* e = getenv("PTTSPP_AAF_CFG"); return e ? atoi(e) : 0; }();  // experiments
  if (cfg == 1) {
    if (C == 32) aa_conv1d_launch_t<32, 8, 7, 2, 4>(d, mBh, mBl, log_alpha, up_f, down_f, rowsA, nst, smem, num_sms, s);
    else aa_conv1d_launch_t<64, 8, 7, 2, 4>(d, mBh, mBl, log_alpha, up_f, down_f, rowsA, nst, smem, num_sms, s);
  } else {
    if (C == 32) aa_conv1d_launch_t<32, 4, 4, 2, 8>(d, mBh, mBl, log_alpha, up_f, down_f, rowsA, nst, smem, num_sms, s);
    else aa_conv1d_launch_t<64, 8, 7, 1, 8>(d, mBh, mBl, log_alpha, up_f, down_f, rowsA, nst, smem, num_sms, s);
  }
}

}  // namespace

bool conv1d_umma_supported(const pttspp_conv1d_desc& d) {
  if (conv1d_umma_c32_ok(d)) return true;
  return d.in_hi && d.in_lo && d.w_hi && d.w_lo && d.Cin % UM_BK == 0 && d.in_stride == 1 && !d.in_len && !d.in_add &&
         d.in_ld % 8 == 0 && d.in_bs % 8 == 0 && aligned16(d.in_hi) && aligned16(d.in_lo) && aligned16(d.w_hi) &&
         aligned16(d.w_lo) && d.w_scale_inv > 0.f;
}

// d2 != nullptr: dual-epilogue launch.  d and d2 must describe the same contraction input (planes, K, dil, pad, Cin,
// row range); the packed weight planes of d2 must directly follow those of d (K == 1), so that one tensor map covers
// the Cout(d) + Cout(d2) rows.
void conv1d_umma_launch(const pttspp_conv1d_desc& d_in, const pttspp_conv1d_desc* d2_in, cudaStream_t s) {
  pttspp_conv1d_desc d = d_in;
  d.acc_scale = d_in.acc_scale * d_in.w_scale_inv;  // undo the power-of-two weight scale on the accumulator
  pttspp_conv1d_desc d2 = d2_in ? *d2_in : d_in;
  const int cout1 = d_in.Cout;
  bool vec = conv_epilogue_vec_ok(d);
  if (d2_in) {
    PT_CHECK(d.K == 1 && d2.K == 1 && d2.Cin == d.Cin && d2.in_hi == d.in_hi && d2.M == d.M && d2.m_begin == d.m_begin &&
                 d2.B == d.B && d2.T_in == d.T_in,
             "conv1d dual launch: the two descriptors must share the contraction input");
    PT_CHECK((const char*)d2.w_hi == (const char*)d.w_hi + (size_t)cout1 * d.Cin * 2 &&
                 (const char*)d2.w_lo == (const char*)d.w_lo + (size_t)cout1 * d.Cin * 2 && d2.w_scale_inv == d.w_scale_inv,
             "conv1d dual launch: weight planes must be contiguous");
    PT_CHECK(cout1 % 128 == 0, "conv1d dual launch: first Cout must be a multiple of the N tile");
    d2.acc_scale = d2_in->acc_scale * d2_in->w_scale_inv;
    vec = vec && conv_epilogue_vec_ok(d2);
    d.Cout = cout1 + d2.Cout;  // contraction-wide column count (TMA rows, tile count)
  }
  const int num_sms = device_num_sms();
  const int total_cout = d.Cout;
  if (d2_in) d.Cout = cout1;  // the epilogue of the first half sees its own column count again
  if (d_in.impl != 3 && conv1d_umma_pair_launch(d, d2, d2_in != nullptr, cout1, total_cout, s, num_sms)) return;
  PT_CHECK(!d.res_hi && !(d2_in && d2.res_hi),
           "conv1d: a residual from operand planes (res_hi) is only supported by the CTA-pair kernel, which this shape "
           "does not qualify for");
  const int n_mt = ceil_div(d.M, UM_BM), n_nt = ceil_div(total_cout, UM_BN);
  const int nslab = d.Cin / UM_BK;
  // weight planes [K*Cout][Cin]
  const uint64_t wdims[2] = {(uint64_t)d.Cin, (uint64_t)d.K * total_cout};
  const uint64_t wstr[1] = {(uint64_t)d.Cin * 2};
  const uint32_t wbox[2] = {UM_BK, UM_BN};
  const CUtensorMap mBh = make_map(d.w_hi, 2, wdims, wstr, wbox);
  const CUtensorMap mBl = make_map(d.w_lo, 2, wdims, wstr, wbox);
  // activation planes [B][T_in][Cin]: dims innermost first
  const uint64_t adims[3] = {(uint64_t)d.Cin, (uint64_t)d.T_in, (uint64_t)d.B};
  const uint64_t astr[2] = {(uint64_t)d.in_ld * 2, (uint64_t)d.in_bs * 2};

  // A-stationary kernel when the activation block of a 128-row unit fits next to >= 2 weight stages and there are
  // enough units to fill the machine; otherwise the streaming kernel.
  const int rowsA = UM_BM + round_up((d.K - 1) * d.dil, 8);
  const size_t a_bytes = (size_t)nslab * 2 * rowsA * 128;
  const size_t bst = 2 * (size_t)UM_BN * 128;
  const size_t cap = 227 * 1024 - 1024 - 512;
  const long long n_units = (long long)n_mt * d.B;
  // The streaming kernel measured faster on every shape of this model (tools/bench_conv.py: the A-stationary kernel
  // is left with two weight stages in flight and becomes latency bound), so the latter is opt-in.
  const bool a_stationary = umma_env().a_stationary;
  if (a_stationary && rowsA <= 256 && a_bytes + 2 * bst <= cap && n_units * 2 >= num_sms) {
    const int nabuf = (2 * a_bytes + 3 * bst <= cap) ? 2 : 1;
    const int nbst = (int)std::min<size_t>(6, (cap - nabuf * a_bytes) / bst);
    const uint32_t abox[3] = {UM_BK, (uint32_t)rowsA, 1};
    const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox);
    const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox);
    const size_t smem = nabuf * a_bytes + nbst * bst + 512 + 1024;
    auto kern = conv1d_umma_as_kernel<UM_BN, 1, false>;
    ensure_smem_optin((const void*)kern, 227 * 1024);
    const int grid = (int)std::min<long long>(n_units, num_sms);
    kern<<<grid, UM_THREADS, smem, s>>>(mAh, mAl, mBh, mBl, d, d2, d2_in ? cout1 : total_cout, total_cout, vec ? 1 : 0,
                                        n_mt, n_nt, (int)n_units, rowsA, nbst, nabuf);
    PT_LAUNCHED();
    return;
  }
  // Narrow outputs (<= 64 columns) with many taps -- BigVGAN's 64-channel k = 11 convs, whose weights do not fit the
  // weight-resident kernel: the streaming kernel re-loads the 32 KB activation tile for every tap (L2 -> SM fill: 528 KB
  // per 128-row tile at k = 11), the A-stationary kernel loads one halo block (47 KB, double buffered) and streams only
  // the 16 KB weight tiles (223 KB per tile).  Same MMA order as the streaming kernel's 64-column tile: same bits.
  // PTTSPP_UMMA_NO_AS64=1 keeps the streaming kernel (A/B measurements).
  static const bool no_as64 = getenv("PTTSPP_UMMA_NO_AS64") != nullptr;
  static const bool as64_single = getenv("PTTSPP_UMMA_AS64_NSUB1") != nullptr;  // A/B: 128-row units
  const size_t bst64 = 2 * (size_t)64 * 128;
  const int rowsA2 = 2 * UM_BM + round_up((d.K - 1) * d.dil, 16);  // 256-row units; two TMA boxes of rowsA2 / 2 rows
  const size_t a_bytes2 = (size_t)nslab * 2 * rowsA2 * 128;
  const bool as64 = !no_as64 && !d2_in && total_cout <= 64 && total_cout % 16 == 0 && d.K >= 5 && d.K * d.Cin / 16 <= 64 &&
                    d_in.impl != 3 && n_units * 2 >= num_sms;
  if (as64 && !as64_single && 2 * a_bytes2 + 3 * bst64 <= cap && n_units >= 2 * num_sms) {
    const int nbst = (int)std::min<size_t>(6, (cap - 2 * a_bytes2) / bst64);
    const uint32_t abox[3] = {UM_BK, (uint32_t)(rowsA2 / 2), 1};
    const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox);
    const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox);
    const uint32_t wbox64[2] = {UM_BK, 64};
    const CUtensorMap mBh64 = make_map(d.w_hi, 2, wdims, wstr, wbox64);
    const CUtensorMap mBl64 = make_map(d.w_lo, 2, wdims, wstr, wbox64);
    const size_t smem = 2 * a_bytes2 + nbst * bst64 + 512 + 1024;
    // (measured: the row-per-lane epilogue wins only when the tile also leaves as operand planes -- 811 -> 639 us -- and
    // loses 7 % on plain / residual epilogues)
    auto kern = (d.out_hi && epilogue_rl_ok(d)) ? conv1d_umma_as_kernel<64, 2, true> : conv1d_umma_as_kernel<64, 2, false>;
    ensure_smem_optin((const void*)kern, 227 * 1024);
    const int n_mt2 = ceil_div(n_mt, 2);
    const long long n_units2 = (long long)n_mt2 * d.B;
    const int grid = (int)std::min<long long>(n_units2, num_sms);
    kern<<<grid, UM_THREADS, smem, s>>>(mAh, mAl, mBh64, mBl64, d, d2, total_cout, total_cout, vec ? 1 : 0, n_mt2, 1,
                                        (int)n_units2, rowsA2, nbst, 2);
    PT_LAUNCHED();
    return;
  }
  if (as64 && rowsA <= 256 && 2 * a_bytes + 3 * bst64 <= cap) {
    const int nbst = (int)std::min<size_t>(6, (cap - 2 * a_bytes) / bst64);
    const uint32_t abox[3] = {UM_BK, (uint32_t)rowsA, 1};
    const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox);
    const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox);
    const uint32_t wbox64[2] = {UM_BK, 64};
    const CUtensorMap mBh64 = make_map(d.w_hi, 2, wdims, wstr, wbox64);
    const CUtensorMap mBl64 = make_map(d.w_lo, 2, wdims, wstr, wbox64);
    const size_t smem = 2 * a_bytes + nbst * bst64 + 512 + 1024;
    auto kern = (d.out_hi && epilogue_rl_ok(d)) ? conv1d_umma_as_kernel<64, 1, true> : conv1d_umma_as_kernel<64, 1, false>;
    ensure_smem_optin((const void*)kern, 227 * 1024);
    const int grid = (int)std::min<long long>(n_units, num_sms);
    kern<<<grid, UM_THREADS, smem, s>>>(mAh, mAl, mBh64, mBl64, d, d2, total_cout, total_cout, vec ? 1 : 0, n_mt, 1,
                                        (int)n_units, rowsA, nbst, 2);
    PT_LAUNCHED();
    return;
  }
  const uint32_t abox[3] = {UM_BK, UM_BM, 1};
  const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox);
  const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox);
  // Long contractions (> 64 tensor-core accumulations per accumulator) with >= 128 output columns also take the chunked
  // kernel: 128-column tiles (half the activation re-reads of the 64-column round-robin variant below) and a bounded
  // truncation bias.  PTTSPP_UMMA_LONGK=narrow keeps the old choice (A/B measurements).
  // (impl 4 never chunks: a conv must give the same bits whichever kernel its batch size selects)
  const bool long_k_chunked = !d2_in && total_cout >= 128 && d.K * d.Cin / 16 > 64 && vec && !umma_env().longk_narrow && d_in.impl != 4;
  if (d_in.impl == 3 || long_k_chunked) {
    // near-fp32 chunked accumulation (see the kernel comment): 4 (slab, tap) iterations = 16 accumulations per chunk
    PT_CHECK(!d2_in, "conv1d: the chunked tcgen05 mode has no dual-epilogue form");
    using SMc = UmmaSmem<UM_BN>;
    const size_t smem = (size_t)UM_STAGES * SMc::STAGE_BYTES + 256 + UM_EPI_WARPS * 2048 + 1024;
    auto kern = conv1d_umma_kernel<UM_BN, UM_STAGES, UM_NACC, 2, false>;
    ensure_smem_optin((const void*)kern, (int)smem);
    const long long n_tiles = (long long)n_mt * n_nt * d.B;
    PT_CHECK(n_tiles < (1ll << 30), "conv1d: too many tiles");
    const int grid = (int)std::min<long long>(n_tiles, num_sms);
    OutMaps om0;
    memset(&om0, 0, sizeof(om0));
    kern<<<grid, UM_THREADS, smem, s>>>(mAh, mAl, mBh, mBl, om0, om0, d, d2, total_cout, vec ? 1 : 0, 0, n_mt, n_nt,
                                        (int)n_tiles, umma_env().chunk_iters);
    PT_LAUNCHED();
    return;
  }
  // epilogue output maps (per descriptor, with its own column count)
  pttspp_conv1d_desc e1 = d, e2 = d2;
  if (d2_in) e1.Cout = cout1;
  int tma_out = 0;
  OutMaps om, om2;
  memset(&om, 0, sizeof(om));
  memset(&om2, 0, sizeof(om2));
  // streaming kernel: the TMA bulk-store epilogue measured ~7 % faster than the coalescing one on the BigVGAN shapes
  // (tools/bench_conv.py); PTTSPP_UMMA_EPI=co selects the latter
  const bool want_co = umma_env().epi == 'c';
  const bool tma_possible = vec && tma_out_ok(e1) && (!d2_in || tma_out_ok(e2));
  if ((want_co || !tma_possible) && epilogue_co_ok(e1) && (!d2_in || epilogue_co_ok(e2))) {
    tma_out = 4;  // coalescing epilogue
  } else if (vec && tma_out_ok(e1)) {
    om = make_out_maps(e1);
    tma_out |= 1;
  }
  if (!(tma_out & 4) && d2_in && vec && tma_out_ok(e2)) {
    om2 = make_out_maps(e2);
    tma_out |= 2;
  }
  if (d2_in) d.Cout = total_cout;  // the streaming kernel's producer derives weight rows from d.Cout
  // Long contractions (many MMAs into one accumulator) and narrow outputs take the 64-column tile with three
  // round-robin main accumulators; everything else the 128-column tile with one.
  const int accumulations = d.K * d.Cin / 16;
  const bool narrow = !d2_in && (total_cout <= 64 || accumulations > 64);
  if (narrow) {
    constexpr int BN = 64, ST = 4;
    const uint32_t wbox64[2] = {UM_BK, BN};
    const CUtensorMap mBh64 = make_map(d.w_hi, 2, wdims, wstr, wbox64);
    const CUtensorMap mBl64 = make_map(d.w_lo, 2, wdims, wstr, wbox64);
    const size_t smem = (size_t)ST * UmmaSmem<BN>::STAGE_BYTES + 256 + UM_EPI_WARPS * 2048 + 1024;
    // one main accumulator + the stacked [W_hi | W_lo] issue (two MMAs per k step) by default: the 64-column tile is
    // bound by the operand fetch, and <= 64 truncating accumulations stay far inside the parity bars (measured);
    // PTTSPP_UMMA_NARROW_NACC=4 restores three round-robin main accumulators (three MMAs per k step)
    static const bool rr = [] { const char* e = getenv("PTTSPP_UMMA_NARROW_NACC"); return e && e[0] == '4'; }();
    auto kern = rr ? ((tma_out & 4) ? conv1d_umma_kernel<BN, ST, 4, 1, false> : conv1d_umma_kernel<BN, ST, 4, 0, false>)
                   : ((tma_out & 4) ? conv1d_umma_kernel<BN, ST, 2, 1, false> : conv1d_umma_kernel<BN, ST, 2, 0, false>);
    ensure_smem_optin((const void*)kern, (int)smem);
    const int nnt = ceil_div(total_cout, BN);
    const long long n_tiles = (long long)n_mt * nnt * d.B;
    PT_CHECK(n_tiles < (1ll << 30), "conv1d: too many tiles");
    const int grid = (int)std::min<long long>(n_tiles, num_sms);
    kern<<<grid, UM_THREADS, smem, s>>>(mAh, mAl, mBh64, mBl64, om, om2, d, d2, total_cout, vec ? 1 : 0, tma_out, n_mt,
                                        nnt, (int)n_tiles, 0);
    PT_LAUNCHED();
    return;
  }
  using SM = UmmaSmem<UM_BN>;
  const size_t smem = (size_t)UM_STAGES * SM::STAGE_BYTES + 256 + UM_EPI_WARPS * 2048 + 1024;
  auto kern = d2_in ? ((tma_out & 4) ? conv1d_umma_kernel<UM_BN, UM_STAGES, UM_NACC, 1, true>
                                     : conv1d_umma_kernel<UM_BN, UM_STAGES, UM_NACC, 0, true>)
                    : ((tma_out & 4) ? conv1d_umma_kernel<UM_BN, UM_STAGES, UM_NACC, 1, false>
                                     : conv1d_umma_kernel<UM_BN, UM_STAGES, UM_NACC, 0, false>);
  ensure_smem_optin((const void*)kern, (int)smem);
  const long long n_tiles = (long long)n_mt * n_nt * d.B;
  PT_CHECK(n_tiles < (1ll << 30), "conv1d: too many tiles");
  const int grid = (int)std::min<long long>(n_tiles, num_sms);
  kern<<<grid, UM_THREADS, smem, s>>>(mAh, mAl, mBh, mBl, om, om2, d, d2, d2_in ? cout1 : total_cout, vec ? 1 : 0, tma_out,
                                      n_mt, n_nt, (int)n_tiles, 0);
  PT_LAUNCHED();
}

void conv1d_umma_cl(const pttspp_conv1d_desc& d, cudaStream_t s) {
  if (conv1d_umma_wres_pair_try(d, s)) return;
  if (conv1d_umma_c32_ok(d)) conv1d_umma_c32_launch(d, s);
  else conv1d_umma_launch(d, nullptr, s);
}

bool aa_conv1d_supported(const pttspp_conv1d_desc& d) { return aa_conv1d_ok(d); }

void aa_conv1d_cl(const pttspp_conv1d_desc& d, const float* log_alpha, const float* up_f, const float* down_f,
                  cudaStream_t s) {
  PT_CHECK(log_alpha && up_f && down_f, "aa_conv1d: null activation parameter");
  PT_CHECK(aa_conv1d_ok(d), "aa_conv1d: unsupported descriptor (32 channels with <= 11 taps or 64 channels with <= 7 taps, "
                            "fp32 input rows, split-fp16 weights, 32-byte aligned epilogue tensors)");
  const double rows = (double)d.B * d.M;
  // the activation's algorithmic traffic is gone; account the conv's flops and the fused pass's bytes (read x + write out)
  ProfScope prof(PROF_CONV_UMMA, s, 2.0 * rows * d.Cout * (double)d.Cin * d.K, 4.0 * rows * (d.Cin + d.Cout));
  aa_conv1d_launch(d, log_alpha, up_f, down_f, s);
}

void conv1d_umma_dual_cl(const pttspp_conv1d_desc& d1, const pttspp_conv1d_desc& d2, cudaStream_t s) {
  PT_CHECK(conv1d_umma_supported(d1) && conv1d_umma_supported(d2), "conv1d dual launch: unsupported descriptor");
  const double flops = 2.0 * d1.B * (double)d1.M * (d1.Cout + d2.Cout) * (double)d1.Cin * d1.K;
  ProfScope prof(PROF_CONV_UMMA, s, flops, 0.0);
  conv1d_umma_launch(d1, &d2, s);
}

void umma_probe(const void* a_half /*[rows][64]*/, int rows, const void* b_half /*[128][64]*/, int row_off, int mode,
                float* out, cudaStream_t s) {
  PT_CHECK(rows >= 144 && row_off >= 0 && row_off + 128 <= 144, "umma_probe: bad geometry");
  const uint64_t adims[2] = {64, (uint64_t)rows}, astr[1] = {128};
  const uint32_t abox[2] = {64, 144};
  const uint64_t bdims[2] = {64, 128}, bstr[1] = {128};
  const uint32_t bbox[2] = {64, 128};
  const CUtensorMap mA = make_map(a_half, 2, adims, astr, abox);
  const CUtensorMap mB = make_map(b_half, 2, bdims, bstr, bbox);
  const size_t smem = 144 * 128 + 128 * 128 + 64 + 1024;
  PT_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, s>>>(mA, mB, row_off, mode, out);
  PT_LAUNCHED();
}

}  // namespace pttspp

extern "C" void pttspp_debug_reload_env(void) {
  std::lock_guard<std::mutex> lk(pttspp::g_env_mutex);
  pttspp::umma_env_storage().load();
}

extern "C" int pttspp_umma_probe(const void* a_half, int rows, const void* b_half, int row_off, int mode, float* out,
                                 pttspp_stream_t stream) {
  PT_API_BEGIN
  pttspp::umma_probe(a_half, rows, b_half, row_off, mode, out, (cudaStream_t)stream);
  PT_API_END
}

extern "C" int pttspp_aa_conv1d_cl(const pttspp_conv1d_desc* d, const float* log_alpha, const float* up_filter,
                                   const float* down_filter, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(d, "null descriptor");
  pttspp::aa_conv1d_cl(*d, log_alpha, up_filter, down_filter, (cudaStream_t)stream);
  PT_API_END
}
