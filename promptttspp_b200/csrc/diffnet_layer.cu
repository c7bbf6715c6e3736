// Fused DiffNet residual layers (promptttspp/modules/denoiser.py:69-83) on tcgen05, one persistent cta_group::2 kernel for
// a whole stack of layers of one diffusion step.
//
// Per layer l and 128-row block of one utterance (y = h + step_emb[l] travels as split-fp16 planes):
//     g|f   = dilated_conv_l(y) + b_d + cond_l                 (K = 3 taps x 256 channels -> 512 interleaved columns)
//     z     = sigmoid(g) * tanh(f)                             (256 channels, never leaves the SM)
//     r|s   = W_o z + b_o                                      (256 -> 512: residual | skip)
//     h'    = (h + r) / sqrt(2);   y' = h' + step_emb[l+1]     (written as the planes of the next layer)
//     skip += s
// A task = (layer, unit); a unit = two 128-row blocks, one per CTA of the pair.  Per task the CTA pair
//   * TMA-loads its activation block once (4 slabs x {hi, lo} x 144 rows incl. the dilation halo) and streams 16 KB weight
//     half-tiles through a ring (each CTA loads half of every 128-column weight tile: cta_group::2);
//   * runs the dilated conv as 4 accumulator tiles (D0..D3, 128 columns = 64 z channels each) through TWO TMEM
//     accumulator buffers; the 16 epilogue warps turn tile j into z slab j and store it back into TENSOR MEMORY as packed
//     fp16 hi/lo operand planes (tcgen05.st): z needs 256 of the 512 TMEM columns and no shared memory at all;
//   * runs the 1x1 output projection with the A operand read from TMEM (tcgen05.mma [d], [a_tmem], b_desc): tiles R0, R1
//     (residual) and S0, S1 (skip) through the same two accumulator buffers.
// All three split-fp16 products (hi*hi + hi*lo + lo*hi) accumulate in ONE fp32 accumulator per tile (<= 144 truncating
// accumulations; the same numerics for every batch size).
//
// Layers are chained inside the kernel without a grid barrier: tasks are dealt round-robin in (layer, unit) order to the
// co-resident clusters; a task of layer l waits (acquire) for the completion flags of units u-1, u, u+1 of layer l-1
// (the dilation halo), then fences the async proxy before its TMA loads.  So the machine is filled at task granularity
// over the whole stack (20 x units tasks) instead of per layer, and one launch replaces 40.
#include "common.h"
#include "conv_epilogue.cuh"
#include "diffnet_layer.h"
#include "umma_ptx.cuh"

namespace pttspp {
namespace {

constexpr int DL_C = 256;                       // residual channels
constexpr int DL_NSLAB = DL_C / 64;             // 64-channel K slabs
constexpr int DL_TAPS = 3;
constexpr int DL_EW = 16;                       // epilogue warps: 4 TMEM lane quarters x 4 column groups of 32
constexpr int DL_THREADS = (DL_EW + 2) * 32;
constexpr int DL_ROWS_A = 144;                  // 128 + 2 * 8 halo rows (dilation <= 8)
constexpr int DL_A_PLANE = DL_ROWS_A * 128;     // one (slab, plane) tile
constexpr int DL_A_BYTES = DL_NSLAB * 2 * DL_A_PLANE;
constexpr int DL_BHALF = 64 * 128;              // one plane of a weight half-tile (64 rows x 64 K)
constexpr int DL_BST = 2 * DL_BHALF;            // hi + lo
constexpr int DL_NB = 5;                        // weight ring stages
constexpr int DL_NBARS = 2 * DL_NSLAB + 2 * DL_NB + 4 + DL_NSLAB;
constexpr size_t DL_SMEM = 1024 + DL_A_BYTES + DL_NB * DL_BST + DL_NBARS * 8 + 16;
constexpr uint32_t DL_Z_COL = 256;              // TMEM columns [256, 512): z operand planes, 64 columns per slab (hi | lo)
static_assert(DL_SMEM <= 227 * 1024, "shared memory budget");

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows per CTA x 16 halves = 8 columns) is read from tensor memory
__device__ __forceinline__ void umma_f16_pair_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// release arrive on the barrier at the same offset in CTA `cta` of the cluster (publishes this warp's tcgen05.st)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
// L2-coherent 256-bit load for data written by other SMs inside this kernel (never served from a stale L1 line)
__device__ __forceinline__ f8 ldcg256(const void* p) {
  f8 r;
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ f8 ldna256(const void* p) {
  f8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until the units u-1, u, u+1 of the previous layer have published their outputs (lanes 0..2 poll one flag each)
__device__ __forceinline__ void wait_prev_layer(const unsigned* done_prev, int unit, int n_units, unsigned target, int lane) {
  if (lane < 3) {
    const int u = unit - 1 + lane;
    if (u >= 0 && u < n_units) {
      long long t0 = 0;
      for (uint32_t spin = 0; ld_acquire_gpu(done_prev + u) < target; ++spin) {
        if ((spin & 0x3FFu) == 0x3FFu) {
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000ll) __trap();  // a protocol bug traps instead of hanging the GPU
        }
      }
    }
  }
  __syncwarp();
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// barrier wait that accounts its duration (profiling runs only)
__device__ __forceinline__ void mbar_wait_warp_t(uint32_t bar, uint32_t parity, unsigned long long* ctr) {
  if (ctr) {
    const long long t0 = clock64();
    mbar_wait_warp(bar, parity);
    if ((threadIdx.x & 31) == 0) *ctr += (unsigned long long)(clock64() - t0);
  } else {
    mbar_wait_warp(bar, parity);
  }
}

struct TaskGeom {
  int layer, unit, b, mt;
  bool blk_ok;
};
__device__ __forceinline__ TaskGeom task_geom(const DiffNetArgs& a, int t, uint32_t rank) {
  TaskGeom g;
  g.layer = a.layer_begin + t / a.n_units;
  g.unit = t - (t / a.n_units) * a.n_units;
  const int q = 2 * g.unit + (int)rank;  // flat (utterance, block) index: the pair's CTAs may work on different utterances
  g.blk_ok = q / a.n_mt < a.B;
  g.b = min(q / a.n_mt, a.B - 1);
  g.mt = g.blk_ok ? q - (q / a.n_mt) * a.n_mt : a.n_mt;  // past the end: a block without rows (TMA zero-fills)
  return g;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DL_THREADS, 1)
diffnet_layers_kernel(const __grid_constant__ CUtensorMap mapY0h, const __grid_constant__ CUtensorMap mapY0l,
                      const __grid_constant__ CUtensorMap mapY1h, const __grid_constant__ CUtensorMap mapY1l,
                      const __grid_constant__ DiffNetArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ring = base + DL_A_BYTES;
  constexpr uint32_t bars_off = DL_A_BYTES + DL_NB * DL_BST;
  const uint32_t bars = base + bars_off;
  auto fullA = [&](int sl) { return bars + sl * 8; };
  auto emptyA = [&](int sl) { return bars + (DL_NSLAB + sl) * 8; };
  auto fullB = [&](int st) { return bars + (2 * DL_NSLAB + st) * 8; };
  auto emptyB = [&](int st) { return bars + (2 * DL_NSLAB + DL_NB + st) * 8; };
  auto tfull_bar = [&](int u) { return bars + (2 * DL_NSLAB + 2 * DL_NB + u) * 8; };
  auto tempty_bar = [&](int u) { return bars + (2 * DL_NSLAB + 2 * DL_NB + 2 + u) * 8; };
  auto zfull = [&](int j) { return bars + (2 * DL_NSLAB + 2 * DL_NB + 4 + j) * 8; };
  const uint32_t tmem_slot = bars + DL_NBARS * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + bars_off + DL_NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_tasks = (a.layer_end - a.layer_begin) * a.n_units;
  const int last_layer = a.n_layers_total - 1;

  if (threadIdx.x == 0) {
    for (int sl = 0; sl < DL_NSLAB; ++sl) {
      mbar_init(fullA(sl), 1);
      mbar_init(emptyA(sl), 1);
      mbar_init(zfull(sl), 2 * DL_EW);  // every epilogue warp of both CTAs stores a part of every z slab
    }
    for (int st = 0; st < DL_NB; ++st) {
      mbar_init(fullB(st), 1);
      mbar_init(emptyB(st), 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(tfull_bar(u), 1);
      mbar_init(tempty_bar(u), 2 * DL_EW);
    }
    fence_barrier_init();
  }
  if (warp == DL_EW && lane == 0) {
    tma_prefetch_desc(&mapY0h);
    tma_prefetch_desc(&mapY0l);
    tma_prefetch_desc(&mapY1h);
    tma_prefetch_desc(&mapY1l);
  }
  if (warp == DL_EW + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == DL_EW) {
    // ================= TMA producer (both CTAs: own activation rows, own half of every weight tile) =================
    uint32_t g = 0;
    int ablk = 0;
    unsigned long long* pc = (a.dbg_prof && rank == 0) ? a.dbg_prof + (size_t)cluster_id * 16 : nullptr;
    for (int t = cluster_id; t < n_tasks; t += n_clusters, ++ablk) {
      const TaskGeom tg = task_geom(a, t, rank);
      const DiffLayerConst* L = a.layers + tg.layer;
      const int dil = L->dil;
      const int row0 = tg.mt * 128 - dil;
      if (tg.layer > a.layer_begin) {
        const long long tw = pc ? clock64() : 0;
        wait_prev_layer(a.done + (size_t)(tg.layer - 1) * a.n_units, tg.unit, a.n_units, a.done_target, lane);
        if (pc && lane == 0) pc[10] += (unsigned long long)(clock64() - tw);
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of other SMs -> this warp's TMA reads
      }
      const CUtensorMap* mYh = (tg.layer & 1) ? &mapY1h : &mapY0h;
      const CUtensorMap* mYl = (tg.layer & 1) ? &mapY1l : &mapY0l;
      const int n_o_begin = (a.dbg_flags & 4) ? 4 : ((tg.layer == last_layer) ? 2 : 0);  // the last layer's residual half has no consumer
      for (int nt = 0; nt < 4; ++nt)
        for (int slab = 0; slab < DL_NSLAB; ++slab) {
          if (nt == 0) {
            mbar_wait_warp_t(emptyA(slab), ((uint32_t)ablk & 1u) ^ 1u, pc ? pc + 8 : nullptr);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(fullA(slab), 4u * DL_A_PLANE);
              tma_load_3d_pair(base + (uint32_t)(2 * slab) * DL_A_PLANE, mYh, fullA(slab), slab * 64, row0, tg.b);
              tma_load_3d_pair(base + (uint32_t)(2 * slab + 1) * DL_A_PLANE, mYl, fullA(slab), slab * 64, row0, tg.b);
            }
            __syncwarp();
          }
          for (int tap = 0; tap < DL_TAPS; ++tap, ++g) {
            const int st = g % DL_NB;
            mbar_wait_warp_t(emptyB(st), ((g / DL_NB) & 1u) ^ 1u, pc ? pc + 9 : nullptr);
            const uint32_t dst = ring + (uint32_t)st * DL_BST;
            const int wrow = tap * (2 * DL_C) + nt * 128 + (int)rank * 64;
            if (elect_one()) {
              if (a.dbg_flags & 256) {  // experiment: half the weight traffic (lo planes not loaded; results invalid)
                if (rank == 0) mbar_expect_tx(fullB(st), (uint32_t)DL_BST);
                tma_load_2d_pair(dst, &L->wd_h, fullB(st), slab * 64, wrow);
              } else {
                if (rank == 0) mbar_expect_tx(fullB(st), 2u * DL_BST);
                tma_load_2d_pair(dst, &L->wd_h, fullB(st), slab * 64, wrow);
                tma_load_2d_pair(dst + DL_BHALF, &L->wd_l, fullB(st), slab * 64, wrow);
              }
            }
            __syncwarp();
          }
        }
      for (int nt = n_o_begin; nt < 4; ++nt)
        for (int slab = 0; slab < DL_NSLAB; ++slab, ++g) {
          const int st = g % DL_NB;
          mbar_wait_warp(emptyB(st), ((g / DL_NB) & 1u) ^ 1u);
          const uint32_t dst = ring + (uint32_t)st * DL_BST;
          const int wrow = nt * 128 + (int)rank * 64;
          if (elect_one()) {
            if (a.dbg_flags & 256) {
              if (rank == 0) mbar_expect_tx(fullB(st), (uint32_t)DL_BST);
              tma_load_2d_pair(dst, &L->wo_h, fullB(st), slab * 64, wrow);
            } else {
              if (rank == 0) mbar_expect_tx(fullB(st), 2u * DL_BST);
              tma_load_2d_pair(dst, &L->wo_h, fullB(st), slab * 64, wrow);
              tma_load_2d_pair(dst + DL_BHALF, &L->wo_l, fullB(st), slab * 64, wrow);
            }
          }
          __syncwarp();
        }
    }
  } else if (warp == DL_EW + 1) {
    // ================= MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) =================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(256, 128);
      const uint64_t desc0 = umma_desc_k_sw128(base);  // descriptors differ only in the start-address field
      uint32_t g = 0;
      int i = 0, ablk = 0;
      unsigned long long* pc = a.dbg_prof ? a.dbg_prof + (size_t)cluster_id * 16 : nullptr;
      const long long t_start = pc ? clock64() : 0;
      for (int t = cluster_id; t < n_tasks; t += n_clusters, ++ablk) {
        const int layer = a.layer_begin + t / a.n_units;
        const int dil = a.layers[layer].dil;
        const int n_o_begin = (a.dbg_flags & 4) ? 4 : ((layer == last_layer) ? 2 : 0);
        const bool one_mma = (a.dbg_flags & 2) != 0;
        // ---- dilated conv: tiles D0..D3 (operands from shared memory) ----
        for (int nt = 0; nt < 4; ++nt, ++i) {
          const int u = i & 1;
          mbar_wait_warp_t(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u, pc ? pc + 0 : nullptr);  // both CTAs have drained accumulator buffer u
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(u * 128);
          uint32_t first = 0;
          for (int slab = 0; slab < DL_NSLAB; ++slab) {
            if (nt == 0) {
              mbar_wait_warp_t(fullA(slab), (uint32_t)ablk & 1u, pc ? pc + 2 : nullptr);
              tc_fence_after();
            }
            for (int tap = 0; tap < DL_TAPS; ++tap, ++g) {
              const int st = g % DL_NB;
              mbar_wait_warp_t(fullB(st), (g / DL_NB) & 1u, pc ? pc + 3 : nullptr);
              tc_fence_after();
              const uint32_t a_off = (uint32_t)(2 * slab) * DL_A_PLANE + (uint32_t)(tap * dil) * 128u;  // taps share the halo tile
              const uint64_t dAh = desc0 + (uint64_t)(a_off >> 4);
              const uint64_t dAl = dAh + (uint64_t)(DL_A_PLANE >> 4);
              const uint64_t dBh = desc0 + (uint64_t)((DL_A_BYTES + (uint32_t)st * DL_BST) >> 4);
              const uint64_t dBl = dBh + (uint64_t)(DL_BHALF >> 4);
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t adv = (uint64_t)(kk * 2);  // 16 halves = 32 bytes inside the swizzle span
                  umma_f16_pair(acc, dAl + adv, dBh + adv, idesc, kk ? 1u : first);
                  if (one_mma) continue;
                  umma_f16_pair(acc, dAh + adv, dBl + adv, idesc, 1u);
                  umma_f16_pair(acc, dAh + adv, dBh + adv, idesc, 1u);
                }
                umma_commit_pair(emptyB(st));
                if (nt == 3 && tap == DL_TAPS - 1) umma_commit_pair(emptyA(slab));  // the next task's block may stream in
                if (slab == DL_NSLAB - 1 && tap == DL_TAPS - 1) umma_commit_pair(tfull_bar(u));
              }
              __syncwarp();
              first = 1u;
            }
          }
        }
        // ---- 1x1 output projection: tiles R0, R1, S0, S1 (A operand = z planes in tensor memory) ----
        for (int nt = n_o_begin; nt < 4; ++nt, ++i) {
          const int u = i & 1;
          mbar_wait_warp_t(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u, pc ? pc + 1 : nullptr);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(u * 128);
          for (int slab = 0; slab < DL_NSLAB; ++slab, ++g) {
            if (nt == n_o_begin) {
              mbar_wait_warp_t(zfull(slab), (uint32_t)ablk & 1u, pc ? pc + 5 : nullptr);  // z slab `slab` of both CTAs is in tensor memory
              tc_fence_after();
            }
            const int st = g % DL_NB;
            mbar_wait_warp_t(fullB(st), (g / DL_NB) & 1u, pc ? pc + 4 : nullptr);
            tc_fence_after();
            const uint64_t dBh = desc0 + (uint64_t)((DL_A_BYTES + (uint32_t)st * DL_BST) >> 4);
            const uint64_t dBl = dBh + (uint64_t)(DL_BHALF >> 4);
            const uint32_t zh = tmem_base + DL_Z_COL + (uint32_t)(slab * 64), zl = zh + 32u;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 2);
                umma_f16_pair_ts(acc, zl + (uint32_t)(kk * 8), dBh + adv, idesc, (slab | kk) ? 1u : 0u);
                if (one_mma) continue;
                umma_f16_pair_ts(acc, zh + (uint32_t)(kk * 8), dBl + adv, idesc, 1u);
                umma_f16_pair_ts(acc, zh + (uint32_t)(kk * 8), dBh + adv, idesc, 1u);
              }
              umma_commit_pair(emptyB(st));
              if (slab == DL_NSLAB - 1) umma_commit_pair(tfull_bar(u));
            }
            __syncwarp();
          }
        }
      }
      if (pc && lane == 0) pc[7] += (unsigned long long)(clock64() - t_start);
    }
  } else {
    // ================= epilogue: warps 0-15 of both CTAs, each CTA drains its own 128 TMEM lanes =================
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float inv_sqrt2 = 0.70710678118654752440f;
    int i = 0;
    unsigned long long* pc = (a.dbg_prof && rank == 0 && warp == 0) ? a.dbg_prof + (size_t)cluster_id * 16 : nullptr;
    auto release_acc = [&](int u) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar(u));
        else mbar_arrive_cluster(tempty_bar(u), 0);
      }
    };
    for (int t = cluster_id; t < n_tasks; t += n_clusters) {
      const TaskGeom tg = task_geom(a, t, rank);
      const DiffLayerConst* L = a.layers + tg.layer;
      const int row = tg.mt * 128 + q * 32 + lane;
      const bool ok = tg.blk_ok && row < a.T && !(a.dbg_flags & 1);
      const int64_t rix = (int64_t)tg.b * a.T + row;  // flat row index of this lane
      const float scale_d = L->scale_d, scale_o = L->scale_o;
      const int n_o_begin = (a.dbg_flags & 4) ? 4 : ((tg.layer == last_layer) ? 2 : 0);
      // ---- D tiles: z = sigmoid(g) * tanh(f) of (acc + bias + cond), stored to tensor memory as operand planes ----
      const float* cond_row = a.cond + (int64_t)tg.layer * a.cond_layer_stride + rix * (2 * DL_C);
#pragma unroll 1
      for (int nt = 0; nt < 4; ++nt, ++i) {
        const int u = i & 1;
        const int col0 = nt * 128 + cg * 32;  // first pre-activation column of this warp
        f8 pre[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ok && !(a.dbg_flags & 8)) pre[k] = (a.dbg_flags & 64) ? ldna256(cond_row + col0 + 8 * k) : ldg256(cond_row + col0 + 8 * k);
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) pre[k].v[e] = 0.f;
          }
        }
        mbar_wait_warp_t(tfull_bar(u), ((uint32_t)i >> 1) & 1u, pc ? pc + 11 : nullptr);
        tc_fence_after();
        uint32_t acc[2][16];
        tmem_ld16_nowait(lane_base + (uint32_t)(u * 128 + cg * 32), acc[0]);
        tmem_ld16_nowait(lane_base + (uint32_t)(u * 128 + cg * 32 + 16), acc[1]);
        tmem_wait_ld();
        release_acc(u);
        uint32_t zh[8], zl[8];
        float zf[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int e0 = 2 * p, e1 = 2 * p + 1;
          const float2 bz = __ldg(reinterpret_cast<const float2*>(L->bias_d + col0 + e0));
          const float gpre = __uint_as_float(acc[e0 >> 4][e0 & 15]) * scale_d + bz.x + pre[e0 >> 3].v[e0 & 7];
          const float fpre = __uint_as_float(acc[e1 >> 4][e1 & 15]) * scale_d + bz.y + pre[e1 >> 3].v[e1 & 7];
          zf[p] = gate_fast(gpre, fpre);
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) split2_f16(zf[2 * m], zf[2 * m + 1], zh[m], zl[m]);
        const uint32_t zaddr = lane_base + DL_Z_COL + (uint32_t)(nt * 64 + cg * 8);
        tmem_st8(zaddr, zh);
        tmem_st8(zaddr + 32u, zl);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(zfull(nt), 0);  // relaxed: tcgen05.wait::st + fence precede it (like tempty)
        if (a.dbg_z && ok) {
          float* dz = a.dbg_z + rix * DL_C + nt * 64 + cg * 16;
          stg256(dz, zf);
          stg256(dz + 8, zf + 8);
        }
      }
      // this warp's own acquire of the previous layer's outputs before it reads them with generic loads
      if (tg.layer > a.layer_begin)
        wait_prev_layer(a.done + (size_t)(tg.layer - 1) * a.n_units, tg.unit, a.n_units, a.done_target, lane);
      const __half* yin_h = (tg.layer & 1) ? a.y_hi[1] : a.y_hi[0];
      const __half* yin_l = (tg.layer & 1) ? a.y_lo[1] : a.y_lo[0];
      __half* yout_h = (tg.layer & 1) ? a.y_hi[0] : a.y_hi[1];
      __half* yout_l = (tg.layer & 1) ? a.y_lo[0] : a.y_lo[1];
      const float* step_cur = a.step_emb + (size_t)tg.layer * DL_C;
      const float* step_next = step_cur + DL_C;
#pragma unroll 1
      for (int nt = n_o_begin; nt < 4; ++nt, ++i) {
        const int u = i & 1;
        const bool is_res = nt < 2;
        const int col0 = (nt & 1) * 128 + cg * 32;  // channel of the residual / skip half
        f8 pre[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int e = 0; e < 8; ++e) pre[k].v[e] = 0.f;
        }
        if (ok && !(a.dbg_flags & 16)) {
          if (is_res) {  // residual from the operand planes: 32 halves of each plane, raw bits
            pre[0] = ldcg256(yin_h + rix * DL_C + col0);
            pre[1] = ldcg256(yin_h + rix * DL_C + col0 + 16);
            pre[2] = ldcg256(yin_l + rix * DL_C + col0);
            pre[3] = ldcg256(yin_l + rix * DL_C + col0 + 16);
          } else if (tg.layer > 0) {  // skip accumulator
#pragma unroll
            for (int k = 0; k < 4; ++k) pre[k] = ldcg256(a.skip + rix * DL_C + col0 + 8 * k);
          }
        }
        mbar_wait_warp_t(tfull_bar(u), ((uint32_t)i >> 1) & 1u, pc ? pc + 12 : nullptr);
        tc_fence_after();
        uint32_t acc[2][16];
        tmem_ld16_nowait(lane_base + (uint32_t)(u * 128 + cg * 32), acc[0]);
        tmem_ld16_nowait(lane_base + (uint32_t)(u * 128 + cg * 32 + 16), acc[1]);
        tmem_wait_ld();
        release_acc(u);
        if (ok && !(a.dbg_flags & 32)) {
          const float* bias = L->bias_o + (is_res ? 0 : DL_C) + col0;
          float v[32];
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 bz = __ldg(reinterpret_cast<const float4*>(bias + 4 * e4));
            v[4 * e4 + 0] = __uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 0) & 15]) * scale_o + bz.x;
            v[4 * e4 + 1] = __uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 1) & 15]) * scale_o + bz.y;
            v[4 * e4 + 2] = __uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 2) & 15]) * scale_o + bz.z;
            v[4 * e4 + 3] = __uint_as_float(acc[(4 * e4) >> 4][(4 * e4 + 3) & 15]) * scale_o + bz.w;
          }
          if (is_res) {
            // h = hi + lo - step_emb[l];  h' = (h + r) / sqrt(2);  y' = h' + step_emb[l+1]
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const uint32_t hb = __float_as_uint(pre[p >> 3].v[p & 7]), lb = __float_as_uint(pre[2 + (p >> 3)].v[p & 7]);
              const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hb));
              const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lb));
              const float2 sc = __ldg(reinterpret_cast<const float2*>(step_cur + col0 + 2 * p));
              const float2 sn = __ldg(reinterpret_cast<const float2*>(step_next + col0 + 2 * p));
              const float h0 = (hf.x + lf.x) - sc.x, h1 = (hf.y + lf.y) - sc.y;
              const float y0 = (h0 + v[2 * p]) * inv_sqrt2 + sn.x, y1 = (h1 + v[2 * p + 1]) * inv_sqrt2 + sn.y;
              split2_f16(y0, y1, hw[p], lw[p]);
            }
            stg256u(yout_h + rix * DL_C + col0, hw);
            stg256u(yout_h + rix * DL_C + col0 + 16, hw + 8);
            stg256u(yout_l + rix * DL_C + col0, lw);
            stg256u(yout_l + rix * DL_C + col0 + 16, lw + 8);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += pre[e >> 3].v[e & 7];
            if (tg.layer == last_layer && a.skip_hi) {  // operand planes of the skip sum for the skip projection
              uint32_t hw[16], lw[16];
#pragma unroll
              for (int p = 0; p < 16; ++p) split2_f16(v[2 * p], v[2 * p + 1], hw[p], lw[p]);
              stg256u(a.skip_hi + rix * DL_C + col0, hw);
              stg256u(a.skip_hi + rix * DL_C + col0 + 16, hw + 8);
              stg256u(a.skip_lo + rix * DL_C + col0, lw);
              stg256u(a.skip_lo + rix * DL_C + col0 + 16, lw + 8);
            } else {
              float* dst = a.skip + rix * DL_C + col0;
#pragma unroll
              for (int k = 0; k < 4; ++k) stg256(dst + 8 * k, v + 8 * k);
            }
          }
        }
      }
      // publish this task: every epilogue warp of both CTAs counts once (release at gpu scope after its stores)
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(a.done + (size_t)tg.layer * a.n_units + tg.unit, 1u);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == DL_EW + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn dl_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
CUtensorMap dl_make_map(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = dl_encode_fn();
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

// ---- host side -------------------------------------------------------------------------------------------------------
DiffNetStack::~DiffNetStack() {
  if (d_layers) cudaFree(d_layers);
}

void DiffNetStack::set_layers(const std::vector<DiffLayerHost>& layers) {
  PT_CHECK(!layers.empty() && layers.size() <= 64, "diffnet: bad layer count");
  std::vector<DiffLayerConst> host(layers.size());
  for (size_t l = 0; l < layers.size(); ++l) {
    const DiffLayerHost& h = layers[l];
    PT_CHECK(h.dil >= 1 && h.dil <= 8, "diffnet: dilation %d unsupported by the fused kernel (halo of 8 rows)", h.dil);
    PT_CHECK(h.wd_hi && h.wd_lo && h.wo_hi && h.wo_lo && h.bias_d && h.bias_o, "diffnet: null layer tensor");
    DiffLayerConst& c = host[l];
    memset(&c, 0, sizeof(c));
    const uint64_t ddims[2] = {(uint64_t)DL_C, (uint64_t)DL_TAPS * 2 * DL_C}, odims[2] = {(uint64_t)DL_C, (uint64_t)2 * DL_C};
    const uint64_t wstr[1] = {(uint64_t)DL_C * 2};
    const uint32_t wbox[2] = {64, 64};
    c.wd_h = dl_make_map(h.wd_hi, 2, ddims, wstr, wbox);
    c.wd_l = dl_make_map(h.wd_lo, 2, ddims, wstr, wbox);
    c.wo_h = dl_make_map(h.wo_hi, 2, odims, wstr, wbox);
    c.wo_l = dl_make_map(h.wo_lo, 2, odims, wstr, wbox);
    c.bias_d = h.bias_d;
    c.bias_o = h.bias_o;
    c.scale_d = h.scale_d;
    c.scale_o = h.scale_o;
    c.dil = h.dil;
  }
  if (d_layers) PT_CUDA(cudaFree(d_layers));
  d_layers = nullptr;
  PT_CUDA(cudaMalloc(&d_layers, host.size() * sizeof(DiffLayerConst)));
  PT_CUDA(cudaMemcpy(d_layers, host.data(), host.size() * sizeof(DiffLayerConst), cudaMemcpyHostToDevice));
  n_layers = (int)layers.size();
  y_key[0] = nullptr;
}

size_t DiffNetStack::flags_bytes(int B, int T) const {
  const int n_mt = ceil_div(T, 128);
  return (size_t)n_layers * (size_t)ceil_div(B * n_mt, 2) * sizeof(unsigned);
}

void DiffNetStack::run(const DiffNetRun& r, cudaStream_t s) {
  PT_CHECK(d_layers && n_layers > 0, "diffnet: layers not set");
  PT_CHECK(r.B >= 1 && r.T >= 1 && r.layer_begin >= 0 && r.layer_begin < r.layer_end && r.layer_end <= n_layers,
           "diffnet: bad geometry");
  PT_CHECK(r.cond && r.step_emb && r.y_hi[0] && r.y_lo[0] && r.y_hi[1] && r.y_lo[1] && r.skip && r.done, "diffnet: null argument");
  auto a32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
  PT_CHECK(a32(r.cond) && a32(r.y_hi[0]) && a32(r.y_lo[0]) && a32(r.y_hi[1]) && a32(r.y_lo[1]) && a32(r.skip) &&
               (!r.skip_hi || (a32(r.skip_hi) && a32(r.skip_lo))),
           "diffnet: tensors must be 32-byte aligned");
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev != setup_dev) {
    PT_CUDA(cudaFuncSetAttribute(diffnet_layers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DL_SMEM));
    int num_sms = 0;
    PT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(DL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = DL_SMEM;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)diffnet_layers_kernel, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    max_clusters = (n > 0) ? std::min(n, num_sms / 2) : num_sms / 2;
    setup_dev = dev;
    y_key[0] = nullptr;
  }
  // activation tensor maps, cached per (buffers, geometry)
  if (y_key[0] != r.y_hi[0] || y_key[1] != r.y_lo[0] || y_key[2] != r.y_hi[1] || y_key[3] != r.y_lo[1] || y_B != r.B || y_T != r.T) {
    const uint64_t adims[3] = {(uint64_t)DL_C, (uint64_t)r.T, (uint64_t)r.B};
    const uint64_t astr[2] = {(uint64_t)DL_C * 2, (uint64_t)r.T * DL_C * 2};
    const uint32_t abox[3] = {64, (uint32_t)DL_ROWS_A, 1};
    const void* ptrs[4] = {r.y_hi[0], r.y_lo[0], r.y_hi[1], r.y_lo[1]};
    for (int k = 0; k < 4; ++k) {
      y_maps[k] = dl_make_map(ptrs[k], 3, adims, astr, abox);
      y_key[k] = ptrs[k];
    }
    y_B = r.B;
    y_T = r.T;
  }
  DiffNetArgs a;
  memset(&a, 0, sizeof(a));
  a.layers = (const DiffLayerConst*)d_layers;
  a.layer_begin = r.layer_begin;
  a.layer_end = r.layer_end;
  a.n_layers_total = n_layers;
  a.B = r.B;
  a.T = r.T;
  a.n_mt = ceil_div(r.T, 128);
  a.n_units = ceil_div(r.B * a.n_mt, 2);
  a.cond = r.cond;
  a.cond_layer_stride = (int64_t)r.B * r.T * 2 * DL_C;
  a.step_emb = r.step_emb;
  for (int k = 0; k < 2; ++k) {
    a.y_hi[k] = (__half*)r.y_hi[k];
    a.y_lo[k] = (__half*)r.y_lo[k];
  }
  a.skip = r.skip;
  a.skip_hi = (__half*)r.skip_hi;
  a.skip_lo = (__half*)r.skip_lo;
  a.done = r.done;
  a.done_target = r.epoch * (2u * DL_EW);
  a.dbg_z = r.dbg_z;
  a.dbg_prof = r.dbg_prof;
  {
    static const int dbg = [] { const char* e = getenv("PTTSPP_DIFFNET_DBG"); return e ? atoi(e) : 0; }();
    a.dbg_flags = dbg;
  }
  const long long n_tasks = (long long)(r.layer_end - r.layer_begin) * a.n_units;
  PT_CHECK(n_tasks < (1ll << 30), "diffnet: too many tasks");
  // all clusters must be co-resident: a task may wait for a task of another cluster (flag protocol)
  const int n_clusters = (int)std::min<long long>(n_tasks, max_clusters);
  void* args[] = {(void*)&y_maps[0], (void*)&y_maps[1], (void*)&y_maps[2], (void*)&y_maps[3], (void*)&a};
  cudaGetLastError();
  PT_CUDA(cudaLaunchKernel((const void*)diffnet_layers_kernel, dim3(2 * n_clusters), dim3(DL_THREADS), args, DL_SMEM, s));
  ++g_launch_count;
}

}  // namespace pttspp

struct pttspp_diffnet {
  pttspp::DiffNetStack stack;
};

extern "C" int pttspp_diffnet_create(const pttspp_diffnet_layer* layers, int n_layers, pttspp_diffnet_t** out) {
  PT_API_BEGIN
  PT_CHECK(layers && out && n_layers >= 1, "null argument");
  std::vector<pttspp::DiffLayerHost> hl(n_layers);
  for (int l = 0; l < n_layers; ++l) {
    const pttspp_diffnet_layer& s = layers[l];
    hl[l] = pttspp::DiffLayerHost{s.wd_hi, s.wd_lo, s.wo_hi, s.wo_lo, s.bias_d, s.bias_o, s.scale_d, s.scale_o, s.dil};
  }
  auto* h = new pttspp_diffnet();
  try {
    h->stack.set_layers(hl);
  } catch (...) {
    delete h;
    throw;
  }
  *out = h;
  PT_API_END
}

extern "C" void pttspp_diffnet_destroy(pttspp_diffnet_t* h) { delete h; }

extern "C" size_t pttspp_diffnet_flags_bytes(const pttspp_diffnet_t* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return h->stack.flags_bytes(B, T);
}

extern "C" int pttspp_diffnet_run(pttspp_diffnet_t* h, const pttspp_diffnet_run_desc* r, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(h && r, "null argument");
  pttspp::DiffNetRun q;
  q.B = r->B; q.T = r->T; q.layer_begin = r->layer_begin; q.layer_end = r->layer_end;
  q.cond = r->cond; q.step_emb = r->step_emb;
  for (int k = 0; k < 2; ++k) { q.y_hi[k] = r->y_hi[k]; q.y_lo[k] = r->y_lo[k]; }
  q.skip = r->skip; q.skip_hi = r->skip_hi; q.skip_lo = r->skip_lo;
  q.done = r->done; q.epoch = r->epoch; q.dbg_z = r->dbg_z; q.dbg_prof = (unsigned long long*)r->dbg_prof;
  const double rows = (double)r->B * r->T * (r->layer_end - r->layer_begin);
  pttspp::ProfScope prof(pttspp::PROF_DIFFNET, (cudaStream_t)stream, rows * 2.0 * (3 * 256 * 512 + 256 * 512), 0.0);
  h->stack.run(q, (cudaStream_t)stream);
  PT_API_END
}
