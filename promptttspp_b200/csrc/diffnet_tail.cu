// The step boundary of the DDPM sampling loop as ONE persistent cta_group::2 tcgen05 kernel (four launches before):
//     p    = relu(W_sp (skip / sqrt(L)) + b_sp)            skip projection, denoiser.py:128-129
//     eps  = W_op p + b_op                                  output projection, :130-131  (80 channels)
//     x'   = coef1 * clamp(c_recip x - c_recipm1 eps) + coef2 x + sigma z        DDPM posterior step, diffusion.py:199-221
//     y    = relu(W_in x' + b_in) + step_emb'[0]            input projection of the NEXT diffusion step, denoiser.py:119-120
// Per unit (two 128-row blocks, one per CTA of the pair): the skip-sum operand planes are TMA-loaded once (4 slabs x
// {hi, lo} x 128 rows), 16 KB weight half-tiles stream through a 5-stage ring, and the three contractions are chained
// THROUGH TENSOR MEMORY: the epilogue of one stores its result as packed split-fp16 operand planes with tcgen05.st and the
// next reads them as its A operand (tcgen05.mma [d], [a_tmem], b_desc) -- p (256 channels) and x' (80 -> 128 channels)
// never touch shared memory or HBM.  HBM traffic per frame: 1 KB of skip planes in, x / noise / x' (3 x 320 B), 1 KB of
// y planes out, against 4 KB + 3 KB of intermediate planes the four separate launches moved.
#include "common.h"
#include "conv_epilogue.cuh"
#include "diffnet_layer.h"
#include "umma_ptx.cuh"

namespace pttspp {
namespace {

constexpr int TL_C = 256;
constexpr int TL_EW = 16;
constexpr int TL_THREADS = (TL_EW + 2) * 32;
constexpr int TL_A_PLANE = 128 * 128;           // one (slab, plane) tile: 128 rows x 128 bytes
constexpr int TL_A_BYTES = 4 * 2 * TL_A_PLANE;  // 4 slabs x {hi, lo}
constexpr int TL_BHALF = 64 * 128;
constexpr int TL_BST = 2 * TL_BHALF;
constexpr int TL_NB = 5;
constexpr int TL_NBARS = 2 * 4 + 2 * TL_NB + 4 + 2;
constexpr size_t TL_SMEM = 1024 + TL_A_BYTES + TL_NB * TL_BST + TL_NBARS * 8 + 16;
constexpr uint32_t TL_P_COL = 256;              // TMEM columns [256, 512): operand planes, 64 columns per 64-channel slab
static_assert(TL_SMEM <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void tl_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tl_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tl_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TL_THREADS, 1)
diffnet_tail_kernel(const __grid_constant__ CUtensorMap mapSh, const __grid_constant__ CUtensorMap mapSl,
                    const __grid_constant__ DiffTailArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base + TL_A_BYTES;
  constexpr uint32_t bars_off = TL_A_BYTES + TL_NB * TL_BST;
  const uint32_t bars = base + bars_off;
  auto fullA = [&](int sl) { return bars + sl * 8; };
  auto emptyA = [&](int sl) { return bars + (4 + sl) * 8; };
  auto fullB = [&](int st) { return bars + (8 + st) * 8; };
  auto emptyB = [&](int st) { return bars + (8 + TL_NB + st) * 8; };
  auto tfull_bar = [&](int u) { return bars + (8 + 2 * TL_NB + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (8 + 2 * TL_NB + 2 + u) * 8; };
  const uint32_t pfull = bars + (8 + 2 * TL_NB + 4) * 8;  // p planes of both CTAs are in tensor memory
  const uint32_t xfull = pfull + 8;                        // x' planes of both CTAs are in tensor memory
  const uint32_t tmem_slot = bars + TL_NBARS * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bars_off + TL_NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const DiffTailConst* cst = a.c;
  const bool do_in = a.step_next != nullptr;

  if (threadIdx.x == 0) {
    for (int sl = 0; sl < 4; ++sl) {
      mbar_init(fullA(sl), 1);
      mbar_init(emptyA(sl), 1);
    }
    for (int st = 0; st < TL_NB; ++st) {
      mbar_init(fullB(st), 1);
      mbar_init(emptyB(st), 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), 2 * TL_EW);
    }
    mbar_init(pfull, 2 * 2 * TL_EW);  // two p tiles x every epilogue warp of both CTAs
    mbar_init(xfull, 2 * TL_EW);
    fence_barrier_init();
  }
  if (warp == TL_EW && lane == 0) {
    tma_prefetch_desc(&mapSh);
    tma_prefetch_desc(&mapSl);
  }
  if (warp == TL_EW + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // flat (utterance, block) index of this CTA's rows in unit `un` (as in diffnet_layers_kernel)
  auto geom = [&](int un, int& b, int& mt, bool& blk_ok) {
    const int q = 2 * un + (int)rank;
    blk_ok = q / a.n_mt < a.B;
    b = min(q / a.n_mt, a.B - 1);
    mt = blk_ok ? q - (q / a.n_mt) * a.n_mt : a.n_mt;
  };

  if (warp == TL_EW) {
    // ================= TMA producer =================
    uint32_t g = 0;
    int ablk = 0;
    auto weight_stage = [&](const CUtensorMap* mh, const CUtensorMap* ml, int kcol, int wrow) {
      const int st = g % TL_NB;
      mbar_wait_warp(emptyB(st), ((g / TL_NB) & 1u) ^ 1u);
      const uint32_t dst = ring + (uint32_t)st * TL_BST;
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(fullB(st), 2u * TL_BST);
        tma_load_2d_pair(dst, mh, fullB(st), kcol, wrow);
        tma_load_2d_pair(dst + TL_BHALF, ml, fullB(st), kcol, wrow);
      }
      __syncwarp();
      ++g;
    };
    for (int un = cluster_id; un < a.n_units; un += n_clusters, ++ablk) {
      int b, mt;
      bool blk_ok;
      geom(un, b, mt, blk_ok);
      for (int nt = 0; nt < 2; ++nt)
        for (int slab = 0; slab < 4; ++slab) {
          if (nt == 0) {
            mbar_wait_warp(emptyA(slab), ((uint32_t)ablk & 1u) ^ 1u);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(fullA(slab), 4u * TL_A_PLANE);
              tma_load_3d_pair(base + (uint32_t)(2 * slab) * TL_A_PLANE, &mapSh, fullA(slab), slab * 64, mt * 128, b);
              tma_load_3d_pair(base + (uint32_t)(2 * slab + 1) * TL_A_PLANE, &mapSl, fullA(slab), slab * 64, mt * 128, b);
            }
            __syncwarp();
          }
          weight_stage(&cst->wsp_h, &cst->wsp_l, slab * 64, nt * 128 + (int)rank * 64);
        }
      for (int slab = 0; slab < 4; ++slab) weight_stage(&cst->wop_h, &cst->wop_l, slab * 64, (int)rank * 64);
      if (do_in)
        for (int nt = 0; nt < 2; ++nt)
          for (int slab = 0; slab < 2; ++slab) weight_stage(&cst->wip_h, &cst->wip_l, slab * 64, nt * 128 + (int)rank * 64);
    }
  } else if (warp == TL_EW + 1) {
    // ================= MMA issuer (leader CTA) =================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(256, 128);
      const uint64_t desc0 = umma_desc_k_sw128(base);
      uint32_t g = 0;
      int i = 0, ablk = 0;
      for (int un = cluster_id; un < a.n_units; un += n_clusters, ++ablk) {
        // ---- p tiles: A = skip planes in shared memory ----
        for (int nt = 0; nt < 2; ++nt, ++i) {
          const int u = i & 1;
          mbar_wait_warp(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(u * 128);
          for (int slab = 0; slab < 4; ++slab, ++g) {
            if (nt == 0) {
              mbar_wait_warp(fullA(slab), (uint32_t)ablk & 1u);
              tc_fence_after();
            }
            const int st = g % TL_NB;
            mbar_wait_warp(fullB(st), (g / TL_NB) & 1u);
            tc_fence_after();
            const uint64_t dAh = desc0 + (uint64_t)(((uint32_t)(2 * slab) * TL_A_PLANE) >> 4);
            const uint64_t dAl = dAh + (uint64_t)(TL_A_PLANE >> 4);
            const uint64_t dBh = desc0 + (uint64_t)((TL_A_BYTES + (uint32_t)st * TL_BST) >> 4);
            const uint64_t dBl = dBh + (uint64_t)(TL_BHALF >> 4);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 2);
                umma_f16_pair(acc, dAl + adv, dBh + adv, idesc, (slab | kk) ? 1u : 0u);
                umma_f16_pair(acc, dAh + adv, dBl + adv, idesc, 1u);
                umma_f16_pair(acc, dAh + adv, dBh + adv, idesc, 1u);
              }
              umma_commit_pair(emptyB(st));
              if (nt == 1) umma_commit_pair(emptyA(slab));  // the next unit's planes may stream in
              if (slab == 3) umma_commit_pair(tfull_bar(u));
            }
            __syncwarp();
          }
        }
        // ---- eps tile: A = p planes in tensor memory (K = 256) ----
        {
          const int u = i & 1;
          mbar_wait_warp(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u);
          mbar_wait_warp(pfull, (uint32_t)ablk & 1u);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(u * 128);
          for (int slab = 0; slab < 4; ++slab, ++g) {
            const int st = g % TL_NB;
            mbar_wait_warp(fullB(st), (g / TL_NB) & 1u);
            tc_fence_after();
            const uint64_t dBh = desc0 + (uint64_t)((TL_A_BYTES + (uint32_t)st * TL_BST) >> 4);
            const uint64_t dBl = dBh + (uint64_t)(TL_BHALF >> 4);
            const uint32_t ph = tmem_base + TL_P_COL + (uint32_t)(slab * 64), pl = ph + 32u;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 2);
                tl_umma_ts(acc, pl + (uint32_t)(kk * 8), dBh + adv, idesc, (slab | kk) ? 1u : 0u);
                tl_umma_ts(acc, ph + (uint32_t)(kk * 8), dBl + adv, idesc, 1u);
                tl_umma_ts(acc, ph + (uint32_t)(kk * 8), dBh + adv, idesc, 1u);
              }
              umma_commit_pair(emptyB(st));
              if (slab == 3) umma_commit_pair(tfull_bar(u));
            }
            __syncwarp();
          }
          ++i;
        }
        // ---- y tiles: A = x' planes in tensor memory (K = 128) ----
        if (do_in) {
          for (int nt = 0; nt < 2; ++nt, ++i) {
            const int u = i & 1;
            mbar_wait_warp(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u);
            if (nt == 0) mbar_wait_warp(xfull, (uint32_t)ablk & 1u);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(u * 128);
            for (int slab = 0; slab < 2; ++slab, ++g) {
              const int st = g % TL_NB;
              mbar_wait_warp(fullB(st), (g / TL_NB) & 1u);
              tc_fence_after();
              const uint64_t dBh = desc0 + (uint64_t)((TL_A_BYTES + (uint32_t)st * TL_BST) >> 4);
              const uint64_t dBl = dBh + (uint64_t)(TL_BHALF >> 4);
              const uint32_t xh = tmem_base + TL_P_COL + (uint32_t)(slab * 64), xl = xh + 32u;
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t adv = (uint64_t)(kk * 2);
                  tl_umma_ts(acc, xl + (uint32_t)(kk * 8), dBh + adv, idesc, (slab | kk) ? 1u : 0u);
                  tl_umma_ts(acc, xh + (uint32_t)(kk * 8), dBl + adv, idesc, 1u);
                  tl_umma_ts(acc, xh + (uint32_t)(kk * 8), dBh + adv, idesc, 1u);
                }
                umma_commit_pair(emptyB(st));
                if (slab == 1) umma_commit_pair(tfull_bar(u));
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else {
    // ================= epilogue: warps 0-15 of both CTAs, each CTA drains its own 128 TMEM lanes =================
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int i = 0;
    auto release_acc = [&](int u) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar(u));
        else mbar_arrive_cluster(tempty_bar(u), 0);
      }
    };
    auto load_acc = [&](int u, int idx, uint32_t (&acc)[2][16]) {
      mbar_wait_warp(tfull_bar(u), ((uint32_t)idx >> 1) & 1u);
      tc_fence_after();
      tmem_ld16_nowait(lane_base + (uint32_t)(u * 128 + cg * 32), acc[0]);
      tmem_ld16_nowait(lane_base + (uint32_t)(u * 128 + cg * 32 + 16), acc[1]);
      tmem_wait_ld();
    };
    // 32 values of this lane -> packed operand planes of 64-channel slab `slab`, channel offset (cg & 1) * 32
    auto store_planes = [&](int slab, const float (&v)[32]) {
      uint32_t hw[16], lw[16];
#pragma unroll
      for (int p = 0; p < 16; ++p) split2_f16(v[2 * p], v[2 * p + 1], hw[p], lw[p]);
      const uint32_t addr = lane_base + TL_P_COL + (uint32_t)(slab * 64 + (cg & 1) * 16);
      tl_tmem_st16(addr, hw);
      tl_tmem_st16(addr + 32u, lw);
      tl_tmem_wait_st();
      tc_fence_before();
      __syncwarp();
    };
    for (int un = cluster_id; un < a.n_units; un += n_clusters) {
      int b, mt;
      bool blk_ok;
      geom(un, b, mt, blk_ok);
      const int row = mt * 128 + q * 32 + lane;
      const bool ok = blk_ok && row < a.T;
      const int64_t rix = (int64_t)b * a.T + row;
      // ---- p tiles ----
#pragma unroll 1
      for (int nt = 0; nt < 2; ++nt, ++i) {
        const int u = i & 1;
        const int col0 = nt * 128 + cg * 32;
        uint32_t acc[2][16];
        load_acc(u, i, acc);
        release_acc(u);
        float v[32];
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 bz = __ldg(reinterpret_cast<const float4*>(cst->bias_sp + col0 + 4 * e4));
          v[4 * e4 + 0] = fmaxf(__uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 0) & 15]) * cst->scale_sp + bz.x, 0.f);
          v[4 * e4 + 1] = fmaxf(__uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 1) & 15]) * cst->scale_sp + bz.y, 0.f);
          v[4 * e4 + 2] = fmaxf(__uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 2) & 15]) * cst->scale_sp + bz.z, 0.f);
          v[4 * e4 + 3] = fmaxf(__uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 3) & 15]) * cst->scale_sp + bz.w, 0.f);
        }
        store_planes(nt * 2 + (cg >> 1), v);
        if (lane == 0) mbar_arrive_cluster(pfull, 0);
      }
      // ---- eps tile -> DDPM update -> x' planes ----
      {
        const int u = i & 1;
        const int m0 = cg * 32;               // first mel channel of this warp
        const int nvalid = min(32, a.M - m0);  // 32, 32, 16, <= 0 for M = 80
        float xv[32], zv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) xv[e] = zv[e] = 0.f;
        if (ok && nvalid > 0) {
          const float* xr = a.x + rix * a.M + m0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (8 * k < nvalid) {
              const f8 t8 = ldg256(xr + 8 * k);
#pragma unroll
              for (int e = 0; e < 8; ++e) xv[8 * k + e] = t8.v[e];
            }
          if (a.z) {
            const float* zr = a.z + ((int64_t)b * a.M + m0) * a.T + row;  // [B][M][T]: lanes = consecutive frames
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e < nvalid) zv[e] = __ldg(zr + (int64_t)e * a.T);
          }
        }
        mbar_wait_warp(tfull_bar(u), ((uint32_t)i >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {  // 16 accumulator columns at a time, x' replaces x in place (register budget)
          uint32_t acc[16];
          tmem_ld16_nowait(lane_base + (uint32_t)(u * 128 + cg * 32 + 16 * hb), acc);
          tmem_wait_ld();
          if (hb == 1) release_acc(u);
#pragma unroll
          for (int e2 = 0; e2 < 16; ++e2) {
            const int e = 16 * hb + e2;
            float r = 0.f;
            if (e < nvalid) {
              const float eps = __uint_as_float(acc[e2]) * cst->scale_op + __ldg(cst->bias_op + m0 + e);
              float x0 = a.c_recip * xv[e] - a.c_recipm1 * eps;
              x0 = fminf(fmaxf(x0, -1.f), 1.f);
              r = (a.coef1 * x0 + a.coef2 * xv[e]) + a.sigma * zv[e];
            }
            xv[e] = r;
          }
        }
        if (ok && nvalid > 0) {
          float* xr = a.x + rix * a.M + m0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (8 * k < nvalid) stg256(xr + 8 * k, xv + 8 * k);
        }
        store_planes(cg >> 1, xv);  // channels >= M are zeros: K is padded to 128
        if (lane == 0) mbar_arrive_cluster(xfull, 0);
        ++i;
      }
      // ---- y tiles: operand planes of the next diffusion step's first layer ----
      if (do_in) {
#pragma unroll 1
        for (int nt = 0; nt < 2; ++nt, ++i) {
          const int u = i & 1;
          const int col0 = nt * 128 + cg * 32;
          uint32_t acc[2][16];
          load_acc(u, i, acc);
          release_acc(u);
          if (ok) {
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const float2 bz = __ldg(reinterpret_cast<const float2*>(cst->bias_ip + col0 + 2 * p));
              const float2 sn = __ldg(reinterpret_cast<const float2*>(a.step_next + col0 + 2 * p));
              const float y0 = fmaxf(__uint_as_float(acc[(2 * p) >> 4][(2 * p) & 15]) * cst->scale_ip + bz.x, 0.f) + sn.x;
              const float y1 = fmaxf(__uint_as_float(acc[(2 * p + 1) >> 4][(2 * p + 1) & 15]) * cst->scale_ip + bz.y, 0.f) + sn.y;
              split2_f16(y0, y1, hw[p], lw[p]);
            }
            stg256u(a.y_hi + rix * TL_C + col0, hw);
            stg256u(a.y_hi + rix * TL_C + col0 + 16, hw + 8);
            stg256u(a.y_lo + rix * TL_C + col0, lw);
            stg256u(a.y_lo + rix * TL_C + col0 + 16, lw + 8);
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == TL_EW + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
CUtensorMap tl_make_map(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  PT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap m;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t gbox[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i], gbox[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return m;
}

}  // namespace

DiffNetTail::~DiffNetTail() {
  if (d_const) cudaFree(d_const);
}

void DiffNetTail::set_weights(const DiffTailHost& h) {
  PT_CHECK(h.wsp_hi && h.wsp_lo && h.wop_hi && h.wop_lo && h.wip_hi && h.wip_lo && h.bias_sp && h.bias_op && h.bias_ip,
           "diffnet tail: null weight tensor");
  PT_CHECK(h.mel >= 1 && h.mel <= 128 && h.mel_pad == 128, "diffnet tail: mel_dim %d (padded %d) unsupported", h.mel, h.mel_pad);
  DiffTailConst c;
  memset(&c, 0, sizeof(c));
  const uint32_t wbox[2] = {64, 64};
  {
    const uint64_t dims[2] = {(uint64_t)TL_C, (uint64_t)TL_C}, str[1] = {(uint64_t)TL_C * 2};
    c.wsp_h = tl_make_map(h.wsp_hi, 2, dims, str, wbox);
    c.wsp_l = tl_make_map(h.wsp_lo, 2, dims, str, wbox);
  }
  {  // [mel rows][256]: rows >= mel of the 128-row tile are zero-filled by TMA
    const uint64_t dims[2] = {(uint64_t)TL_C, (uint64_t)h.mel}, str[1] = {(uint64_t)TL_C * 2};
    c.wop_h = tl_make_map(h.wop_hi, 2, dims, str, wbox);
    c.wop_l = tl_make_map(h.wop_lo, 2, dims, str, wbox);
  }
  {  // [256 rows][mel_pad]
    const uint64_t dims[2] = {(uint64_t)h.mel_pad, (uint64_t)TL_C}, str[1] = {(uint64_t)h.mel_pad * 2};
    c.wip_h = tl_make_map(h.wip_hi, 2, dims, str, wbox);
    c.wip_l = tl_make_map(h.wip_lo, 2, dims, str, wbox);
  }
  c.bias_sp = h.bias_sp; c.bias_op = h.bias_op; c.bias_ip = h.bias_ip;
  c.scale_sp = h.scale_sp; c.scale_op = h.scale_op; c.scale_ip = h.scale_ip;
  if (!d_const) PT_CUDA(cudaMalloc(&d_const, sizeof(DiffTailConst)));
  PT_CUDA(cudaMemcpy(d_const, &c, sizeof(c), cudaMemcpyHostToDevice));
  mel = h.mel;
  s_key[0] = nullptr;
}

void DiffNetTail::run(const DiffTailRun& r, cudaStream_t s) {
  PT_CHECK(d_const, "diffnet tail: weights not set");
  PT_CHECK(r.B >= 1 && r.T >= 1 && r.skip_hi && r.skip_lo && r.x, "diffnet tail: bad argument");
  PT_CHECK(!r.step_next || (r.y_hi && r.y_lo), "diffnet tail: the input projection needs output planes");
  auto a32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
  PT_CHECK(a32(r.skip_hi) && a32(r.skip_lo) && a32(r.x) && (!r.y_hi || (a32(r.y_hi) && a32(r.y_lo))) && (mel % 8 == 0),
           "diffnet tail: tensors must be 32-byte aligned (mel %% 8 == 0)");
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev != setup_dev) {
    PT_CUDA(cudaFuncSetAttribute(diffnet_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_SMEM));
    int num_sms = 0;
    PT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    max_clusters = num_sms / 2;
    setup_dev = dev;
    s_key[0] = nullptr;
  }
  if (s_key[0] != r.skip_hi || s_key[1] != r.skip_lo || s_B != r.B || s_T != r.T) {
    const uint64_t adims[3] = {(uint64_t)TL_C, (uint64_t)r.T, (uint64_t)r.B};
    const uint64_t astr[2] = {(uint64_t)TL_C * 2, (uint64_t)r.T * TL_C * 2};
    const uint32_t abox[3] = {64, 128, 1};
    s_maps[0] = tl_make_map(r.skip_hi, 3, adims, astr, abox);
    s_maps[1] = tl_make_map(r.skip_lo, 3, adims, astr, abox);
    s_key[0] = r.skip_hi; s_key[1] = r.skip_lo; s_B = r.B; s_T = r.T;
  }
  DiffTailArgs a;
  memset(&a, 0, sizeof(a));
  a.c = (const DiffTailConst*)d_const;
  a.B = r.B; a.T = r.T; a.M = mel;
  a.n_mt = ceil_div(r.T, 128);
  a.n_units = ceil_div(r.B * a.n_mt, 2);
  a.x = r.x; a.z = r.z;
  a.c_recip = r.c_recip; a.c_recipm1 = r.c_recipm1; a.coef1 = r.coef1; a.coef2 = r.coef2; a.sigma = r.sigma;
  a.step_next = r.step_next;
  a.y_hi = (__half*)r.y_hi; a.y_lo = (__half*)r.y_lo;
  const int n_clusters = std::min(a.n_units, max_clusters);
  void* args[] = {(void*)&s_maps[0], (void*)&s_maps[1], (void*)&a};
  cudaGetLastError();
  PT_CUDA(cudaLaunchKernel((const void*)diffnet_tail_kernel, dim3(2 * n_clusters), dim3(TL_THREADS), args, TL_SMEM, s));
  ++g_launch_count;
}

}  // namespace pttspp
