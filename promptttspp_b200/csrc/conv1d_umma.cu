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

#include "common.h"
#include "conv_epilogue.cuh"

namespace pttspp {
namespace {

constexpr int UM_BM = 128;
constexpr int UM_BK = 64;  // halves per slab row = 128 bytes = one swizzle span
constexpr int UM_THREADS = 320;
constexpr int UM_NACC = 2;  // TMEM accumulators per tile buffer: main (hi*hi) and cross-term (hi*lo + lo*hi)

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN>
struct UmmaSmem {
  static constexpr int A_BYTES = UM_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent, warp-specialised: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The TMA and MMA warps
// run ahead into the next tile while the eight epilogue warps drain the previous accumulator: TMEM holds two
// accumulator buffers of (main | cross-term) x BN columns.
template <int BN, int STAGES>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv1d_umma_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                   const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                   const pttspp_conv1d_desc d, const int vec_ok, const int n_mt, const int n_nt, const int n_tiles) {
  using SM = UmmaSmem<BN>;
  constexpr uint32_t TMEM_COLS = 2 * UM_NACC * BN;  // 2 buffers x (main, cross)
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
      mbar_init(tempty_bar(u), 8);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mapAh);
    tma_prefetch_desc(&mapAl);
    tma_prefetch_desc(&mapBh);
    tma_prefetch_desc(&mapBl);
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 8) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t g = 0;  // ring position, continues across tiles
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % n_nt, mt = (tile / n_nt) % n_mt, b = tile / (n_nt * n_mt);
        const int m0 = d.m_begin + mt * UM_BM, n0 = nt * BN;
        for (int it = 0; it < n_iter; ++it, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const int slab = it / d.K, tap = it % d.K;  // taps innermost: the shifted row windows overlap in L2
          const uint32_t st = base + s * SM::STAGE_BYTES;
          mbar_expect_tx(full_bar(s), SM::STAGE_BYTES);
          const int row = m0 + tap * d.dil - d.pad;
          tma_load_3d(st, &mapAh, full_bar(s), slab * UM_BK, row, b);
          tma_load_3d(st + SM::A_BYTES, &mapAl, full_bar(s), slab * UM_BK, row, b);
          tma_load_2d(st + 2 * SM::A_BYTES, &mapBh, full_bar(s), slab * UM_BK, tap * d.Cout + n0);
          tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES, &mapBl, full_bar(s), slab * UM_BK, tap * d.Cout + n0);
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(UM_BM, BN);
      uint32_t g = 0;
      int i = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
        const int u = i & 1;
        mbar_wait(tempty_bar(u), (((uint32_t)i >> 1) & 1u) ^ 1u);  // epilogue has drained this buffer
        tc_fence_after();
        // The tensor core truncates when it adds into the fp32 accumulator, so the error grows linearly with the
        // number of accumulations into one accumulator: the 2^-11 smaller cross terms (hi*lo, lo*hi) get their
        // own accumulator, the epilogue adds the two in round-to-nearest fp32.
        const uint32_t acc_main = tmem_base + (uint32_t)(u * UM_NACC * BN);
        const uint32_t acc_cross = acc_main + (uint32_t)BN;
        for (int it = 0; it < n_iter; ++it, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t st = base + s * SM::STAGE_BYTES;
          const uint64_t dAh = umma_desc_k_sw128(st);
          const uint64_t dAl = umma_desc_k_sw128(st + SM::A_BYTES);
          const uint64_t dBh = umma_desc_k_sw128(st + 2 * SM::A_BYTES);
          const uint64_t dBl = umma_desc_k_sw128(st + 2 * SM::A_BYTES + SM::B_BYTES);
#pragma unroll
          for (int kk = 0; kk < UM_BK / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // 16 halves = 32 bytes along K inside the swizzle span
            const uint32_t first = (it | kk) != 0;
            umma_f16(acc_cross, dAl + adv, dBh + adv, idesc, first);
            umma_f16(acc_cross, dAh + adv, dBl + adv, idesc, 1u);
            umma_f16(acc_main, dAh + adv, dBh + adv, idesc, first);
          }
          umma_commit(empty_bar(s));  // frees the stage once these MMAs have read it
        }
        umma_commit(tfull_bar(u));  // accumulator buffer u complete
      }
    }
  } else {
    // ================= epilogue: warps 0-7 =================
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int hsel = warp >> 2;  // column half
    int i = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
      const int nt = tile % n_nt, mt = (tile / n_nt) % n_mt, b = tile / (n_nt * n_mt);
      const int m0 = d.m_begin + mt * UM_BM, n0 = nt * BN;
      const int u = i & 1;
      mbar_wait(tfull_bar(u), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const int row = m * d.out_mul + d.out_off;
      const bool row_ok = (m < d.m_begin + d.M) && row >= 0 && row < d.T_out;
      float mask = 1.f;
      if (d.out_len && row_ok) mask = ((long long)row < (long long)d.out_len[b]) ? 1.f : 0.f;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * UM_NACC * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 16) {
        const int col0 = hsel * (BN / 2) + c;
        float v[16], t[16];
        tmem_ld16(tbase + (uint32_t)col0, v);         // warp-collective: no divergence before these
        tmem_ld16(tbase + (uint32_t)(BN + col0), t);  // cross terms
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] += t[e];
        if (row_ok) {
          if (vec_ok && n0 + col0 + 16 <= d.Cout) {
            conv_epilogue16_vec(d, b, row, mask, n0 + col0, v);
          } else {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              const float a4[4] = {v[gq * 4 + 0], v[gq * 4 + 1], v[gq * 4 + 2], v[gq * 4 + 3]};
              conv_epilogue4(d, b, row, mask, n0 + col0 + gq * 4, a4);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(u));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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

CUtensorMap make_map(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  PT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap m;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t gbox[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) gdim[i] = dims[i], gbox[i] = box[i];
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rank %d, dims %llu x %llu)", (int)r, rank,
           (unsigned long long)dims[0], (unsigned long long)dims[1]);
  return m;
}

constexpr int UM_BN = 128;
constexpr int UM_STAGES = 3;

}  // namespace

bool conv1d_umma_supported(const pttspp_conv1d_desc& d) {
  return d.in_hi && d.in_lo && d.w_hi && d.w_lo && d.Cin % UM_BK == 0 && d.in_stride == 1 && !d.in_len && !d.in_add &&
         d.in_ld % 8 == 0 && d.in_bs % 8 == 0 && aligned16(d.in_hi) && aligned16(d.in_lo) && aligned16(d.w_hi) &&
         aligned16(d.w_lo) && d.w_scale_inv > 0.f;
}

void conv1d_umma_cl(const pttspp_conv1d_desc& d_in, cudaStream_t s) {
  pttspp_conv1d_desc d = d_in;
  d.acc_scale = d_in.acc_scale * d_in.w_scale_inv;  // undo the power-of-two weight scale on the accumulator
  // activation planes [B][T_in][Cin]: dims innermost first
  const uint64_t adims[3] = {(uint64_t)d.Cin, (uint64_t)d.T_in, (uint64_t)d.B};
  const uint64_t astr[2] = {(uint64_t)d.in_ld * 2, (uint64_t)d.in_bs * 2};
  const uint32_t abox[3] = {UM_BK, UM_BM, 1};
  // weight planes [K*Cout][Cin]
  const uint64_t wdims[2] = {(uint64_t)d.Cin, (uint64_t)d.K * d.Cout};
  const uint64_t wstr[1] = {(uint64_t)d.Cin * 2};
  const uint32_t wbox[2] = {UM_BK, UM_BN};
  const CUtensorMap mAh = make_map(d.in_hi, 3, adims, astr, abox);
  const CUtensorMap mAl = make_map(d.in_lo, 3, adims, astr, abox);
  const CUtensorMap mBh = make_map(d.w_hi, 2, wdims, wstr, wbox);
  const CUtensorMap mBl = make_map(d.w_lo, 2, wdims, wstr, wbox);
  using SM = UmmaSmem<UM_BN>;
  const size_t smem = (size_t)UM_STAGES * SM::STAGE_BYTES + 256 + 1024;
  auto kern = conv1d_umma_kernel<UM_BN, UM_STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    PT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PT_CUDA(cudaGetDevice(&dev));
    PT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_mt = ceil_div(d.M, UM_BM), n_nt = ceil_div(d.Cout, UM_BN);
  const long long n_tiles = (long long)n_mt * n_nt * d.B;
  PT_CHECK(n_tiles < (1ll << 30), "conv1d: too many tiles");
  const int grid = (int)std::min<long long>(n_tiles, num_sms);
  kern<<<grid, UM_THREADS, smem, s>>>(mAh, mAl, mBh, mBl, d, conv_epilogue_vec_ok(d) ? 1 : 0, n_mt, n_nt, (int)n_tiles);
  PT_LAUNCHED();
}

}  // namespace pttspp
