// Fused relative-position self-attention of the Conformer text encoder on tcgen05 (esp/transformer/attention.py:142-206
// legacy, :237-305 new): ONE kernel per call computes, per (utterance, head, block of 128 queries),
//     S   = (q + u) K^T                        tcgen05.mma, accumulator in tensor memory columns [0, NK)
//     X   = (q + v) P^T  (window of P rows)    tcgen05.mma, accumulator in tensor memory columns [256, 256 + NW)
//     bd' = rel_shift(X)                       X leaves tensor memory ONCE into a shared-memory tile; the shift (incl. the
//                                              legacy variant's wrapped upper triangle, which reads query row i + 1) is a
//                                              per-thread gather from that tile -- the B*H*T*Tp matrix never exists in HBM
//     P   = softmax((S + bd') / sqrt(d_k))     in registers, thread = query row; exp() values go back to tensor memory as
//                                              split-fp16 operand planes (A operand of the last product)
//     O   = P V                                tcgen05.mma with the A operand read from tensor memory, / row sum
// Operands are converted from the fp32 activations to split-fp16 planes (hi = fp16(v), lo = fp16(v - hi)) on the way into
// shared memory, in the canonical K-major SWIZZLE_128B layout; every product runs as hi*hi + hi*lo + lo*hi (<= 48
// tensor-core accumulations per accumulator: the fp32 error class, the scores feed the integer duration rounding).
// Shapes: d_k = 128, T <= 256 (the text side of every shipped config / benchmark); other shapes keep the CUDA-core
// kernels of attention.cu.
#include "common.h"
#include "umma_ptx.cuh"

namespace pttspp {
namespace {

constexpr int AT_DK = 128;
constexpr int AT_THREADS = 256;
constexpr int AT_MAXT = 256;
constexpr int AT_Q_BYTES = 4 * 128 * 128;        // 2 slabs x {hi, lo} x 128 rows x 128 bytes
constexpr int AT_KP_BYTES = 4 * AT_MAXT * 128;   // K rows / P window rows / V^T key slabs
constexpr int AT_XP = 258;                       // pitch of the staged X tile: (XP - 1) % 32 == 1 makes the skewed gather
                                                 // bank-conflict free
constexpr int AT_XS_BYTES = 129 * AT_XP * 4;     // 128 query rows + the row after the block (legacy upper triangle)
constexpr int AT_EXTRA = ((AT_XS_BYTES - AT_KP_BYTES + 1023) / 1024) * 1024;
constexpr size_t AT_SMEM = 1024 + AT_Q_BYTES + AT_KP_BYTES + AT_EXTRA + 4 * 128 * 4 + 64;
static_assert(AT_SMEM <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8_(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// fp32 rows -> split-fp16 planes in the K-major SWIZZLE_128B layout: `rows` rows of 128 values (2 slabs of 64), plane
// (slab, hi|lo) at dst + (slab * 2 + plane) * rows * 128.  src_row(r) returns the row's first value or nullptr (zero row).
template <typename RowFn>
__device__ __forceinline__ void load_rows_kmajor(uint8_t* dst, int rows, RowFn src_row, const float* bias) {
  for (int idx = threadIdx.x; idx < rows * 16; idx += AT_THREADS) {
    const int r = idx >> 4, c8 = idx & 15;
    const float* src = src_row(r);
    float v[8];
    if (src) {
      const float4 a = *reinterpret_cast<const float4*>(src + c8 * 8);
      const float4 b = *reinterpret_cast<const float4*>(src + c8 * 8 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      if (bias) {
        const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
        v[0] += ba.x; v[1] += ba.y; v[2] += ba.z; v[3] += ba.w; v[4] += bb.x; v[5] += bb.y; v[6] += bb.z; v[7] += bb.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2_f16(v[2 * e], v[2 * e + 1], hw[e], lw[e]);
    const int slab = c8 >> 3, chunk = c8 & 7;
    const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((chunk ^ (r & 7)) * 16);
    uint8_t* ph = dst + (size_t)(slab * 2) * rows * 128 + off;
    *reinterpret_cast<uint4*>(ph) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(ph + (size_t)rows * 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// grid (query blocks, H, B), 256 threads: thread = (query row r = 32 * (warp % 4) + lane, column half ch = warp / 4)
__global__ void __launch_bounds__(AT_THREADS, 1)
relpos_attention_umma_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                             const float* __restrict__ p, const float* __restrict__ bias_u,
                             const float* __restrict__ bias_v, const int64_t* __restrict__ lens, int T, int Tp, int H, int ld,
                             int legacy, float scale, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* Qs = gbase;                              // Qu, then Qv planes
  uint8_t* KPs = gbase + AT_Q_BYTES;                // K planes -> P window planes -> staged X (fp32) -> V^T planes
  float* Xs = reinterpret_cast<float*>(KPs);
  float* red = reinterpret_cast<float*>(gbase + AT_Q_BYTES + AT_KP_BYTES + AT_EXTRA);  // [2][2][128]: max, sum per half
  const uint32_t bar = base + AT_Q_BYTES + AT_KP_BYTES + AT_EXTRA + 4 * 128 * 4;
  const uint32_t tmem_slot = bar + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + AT_Q_BYTES + AT_KP_BYTES + AT_EXTRA + 4 * 128 * 4 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = (warp & 3) * 32 + lane, ch = warp >> 2;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * 128;
  const int i = i0 + r;
  const int HD = H * AT_DK;
  long long l64 = lens ? lens[b] : (long long)T;
  const int len = (int)(l64 < (long long)T ? (l64 < 0 ? 0 : l64) : (long long)T);
  const int NK = (T + 15) & ~15;            // key columns (multiple of the MMA N granularity)
  const int n_chunks = NK >> 4;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  const uint64_t descQ = umma_desc_k_sw128(base);
  const uint64_t descKP = umma_desc_k_sw128(base + AT_Q_BYTES);

  const float* qb = q + (int64_t)b * T * ld + h * AT_DK;
  const float* kb_ = k + (int64_t)b * T * ld + h * AT_DK;
  const float* vb = v + (int64_t)b * T * ld + h * AT_DK;
  const float* pb = p + h * AT_DK;

  // A (128 x 128 halves per plane) x B (N rows) -> D columns [dcol, dcol + N): 2 slabs x 4 K steps x 3 products
  auto issue_ss = [&](uint32_t dcol, int N) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    for (int s = 0; s < 2; ++s) {
      const uint64_t dAh = descQ + (uint64_t)(((uint32_t)(s * 2) * 128u * 128u) >> 4);
      const uint64_t dAl = dAh + (uint64_t)((128u * 128u) >> 4);
      const uint64_t dBh = descKP + (uint64_t)(((uint32_t)(s * 2) * (uint32_t)N * 128u) >> 4);
      const uint64_t dBl = dBh + (uint64_t)(((uint32_t)N * 128u) >> 4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t adv = (uint64_t)(kk * 2);
        umma_f16(tmem_base + dcol, dAl + adv, dBh + adv, idesc, (s | kk) ? 1u : 0u);
        umma_f16(tmem_base + dcol, dAh + adv, dBl + adv, idesc, 1u);
        umma_f16(tmem_base + dcol, dAh + adv, dBh + adv, idesc, 1u);
      }
    }
    umma_commit(bar);
  };
  auto wait_mma = [&]() {
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
  };
  auto publish_smem = [&]() {  // generic-proxy shared-memory writes -> visible to the tensor core's async-proxy reads;
    tc_fence_before();         // this thread's tensor-memory accesses are ordered before the MMAs issued after the sync
    fence_proxy_async();
    __syncthreads();
  };

  // ---- S = (q + u) K^T --------------------------------------------------------------------------------------------
  load_rows_kmajor(Qs, 128, [&](int rr) { return (i0 + rr < T) ? qb + (int64_t)(i0 + rr) * ld : nullptr; },
                   bias_u + h * AT_DK);
  load_rows_kmajor(KPs, NK, [&](int rr) { return (rr < T) ? kb_ + (int64_t)rr * ld : nullptr; }, nullptr);
  publish_smem();
  if (threadIdx.x == 0) {
    tc_fence_after();
    issue_ss(0u, NK);
  }
  wait_mma();

  // ---- X windows, rel_shift, scaled + masked scores back into tensor memory ------------------------------------------
  load_rows_kmajor(Qs, 128, [&](int rr) { return (i0 + rr < T) ? qb + (int64_t)(i0 + rr) * ld : nullptr; },
                   bias_v + h * AT_DK);
  const int n_win = legacy ? 1 : ((NK + 127) >> 7);
  float mx = -INFINITY;
  for (int win = 0; win < n_win; ++win) {
    const int k0 = win * 128;
    const int w0 = legacy ? 0 : (T - 1 - i0 - 127 + k0);   // first P row of the window
    const int NW = legacy ? NK : 256;
    if (win > 0) __syncthreads();                           // every thread has finished gathering from the previous tile
    load_rows_kmajor(KPs, NW, [&](int rr) { const int c = w0 + rr; return (c >= 0 && c < Tp) ? pb + (int64_t)c * HD : nullptr; },
                     nullptr);
    publish_smem();
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_ss(256u, NW);
    }
    wait_mma();
    // stage X: thread = (row, column half)
    for (int c16 = ch; c16 < (NW >> 4); c16 += 2) {
      float xv[16];
      tmem_ld16(lane_base + 256u + (uint32_t)(c16 * 16), xv);
#pragma unroll
      for (int e = 0; e < 16; ++e) Xs[r * AT_XP + c16 * 16 + e] = xv[e];
    }
    if (legacy && i0 + 128 < T) {
      // the row after the block, (q_{i0+128} + v) . p_c, on the CUDA cores (one thread per column)
      for (int c = threadIdx.x; c < T; c += AT_THREADS) {
        const float* qr = qb + (int64_t)(i0 + 128) * ld;
        const float* pr = pb + (int64_t)c * HD;
        const float* bv = bias_v + h * AT_DK;
        float acc = 0.f;
#pragma unroll 4
        for (int dd = 0; dd < AT_DK; dd += 4) {
          const float4 qq = *reinterpret_cast<const float4*>(qr + dd);
          const float4 pp = *reinterpret_cast<const float4*>(pr + dd);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bv + dd));
          acc = fmaf(qq.x + bb.x, pp.x, acc);
          acc = fmaf(qq.y + bb.y, pp.y, acc);
          acc = fmaf(qq.z + bb.z, pp.z, acc);
          acc = fmaf(qq.w + bb.w, pp.w, acc);
        }
        Xs[128 * AT_XP + c] = acc;
      }
    }
    __syncthreads();
    const int c_begin = legacy ? 0 : (k0 >> 4), c_end = legacy ? n_chunks : min(n_chunks, (k0 + 128) >> 4);
    for (int c16 = c_begin + ch; c16 < c_end; c16 += 2) {
      float sv[16];
      tmem_ld16(lane_base + (uint32_t)(c16 * 16), sv);
      uint32_t sw[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int j = c16 * 16 + e;
        float bd = 0.f;
        if (i < T && j < T) {
          if (!legacy) bd = Xs[r * AT_XP + (127 - r) + (j - k0)];
          else if (j <= i) bd = Xs[r * AT_XP + (T - 1 - i + j)];
          else if (j > i + 1) bd = Xs[(r + 1) * AT_XP + (j - i - 2)];
        }
        float s = (sv[e] + bd) * scale;
        if (j >= len) s = -INFINITY;
        mx = fmaxf(mx, s);
        sw[e] = __float_as_uint(s);
      }
      tmem_st16(lane_base + (uint32_t)(c16 * 16), sw);
    }
  }
  tmem_wait_st_();
  red[ch * 128 + r] = mx;
  tc_fence_before();
  __syncthreads();  // scores complete; the X tile is dead
  tc_fence_after();

  // ---- V^T planes (B operand of the last product: rows = d, K = keys) ------------------------------------------------
  for (int idx = threadIdx.x; idx < 128 * (NK >> 3); idx += AT_THREADS) {
    const int dd = idx & 127, j8 = idx >> 7;
    float vv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = j8 * 8 + e;
      vv[e] = (j < T) ? vb[(int64_t)j * ld + dd] : 0.f;
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2_f16(vv[2 * e], vv[2 * e + 1], hw[e], lw[e]);
    const int slab = j8 >> 3, chunk = j8 & 7;
    const uint32_t off = (uint32_t)(dd >> 3) * 1024u + (uint32_t)(dd & 7) * 128u + (uint32_t)((chunk ^ (dd & 7)) * 16);
    uint8_t* ph = KPs + (size_t)(slab * 2) * 128 * 128 + off;
    *reinterpret_cast<uint4*>(ph) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(ph + 128 * 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }

  // ---- softmax numerators -> split-fp16 operand planes in tensor memory ----------------------------------------------
  float m = fmaxf(red[r], red[128 + r]);
  if (m == -INFINITY) m = 0.f;  // no valid key (empty utterance): every exp() below is 0
  float sum = 0.f;
  for (int c16 = ch; c16 < n_chunks; c16 += 2) {
    float sv[16];
    tmem_ld16(lane_base + (uint32_t)(c16 * 16), sv);
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float e0 = expf(sv[2 * e] - m), e1 = expf(sv[2 * e + 1] - m);
      sum += e0;
      sum += e1;
      split2_f16(e0, e1, hw[e], lw[e]);
    }
    tmem_st8_(lane_base + 256u + (uint32_t)(c16 * 8), hw);
    tmem_st8_(lane_base + 384u + (uint32_t)(c16 * 8), lw);
  }
  tmem_wait_st_();
  red[256 + ch * 128 + r] = sum;
  tc_fence_before();
  publish_smem();

  // ---- O = P V -----------------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, AT_DK);
    for (int ks = 0; ks < n_chunks; ++ks) {
      const uint64_t dBh = descKP + (uint64_t)((((uint32_t)((ks >> 2) * 2) * 128u * 128u) >> 4) + (uint32_t)((ks & 3) * 2));
      const uint64_t dBl = dBh + (uint64_t)((128u * 128u) >> 4);
      const uint32_t ah = tmem_base + 256u + (uint32_t)(ks * 8), al = tmem_base + 384u + (uint32_t)(ks * 8);
      umma_f16_ts(tmem_base, al, dBh, idesc, ks ? 1u : 0u);
      umma_f16_ts(tmem_base, ah, dBl, idesc, 1u);
      umma_f16_ts(tmem_base, ah, dBh, idesc, 1u);
    }
    umma_commit(bar);
  }
  wait_mma();
  const float inv = 1.f / (red[256 + r] + red[256 + 128 + r]);
  {
    float* orow = out + ((int64_t)b * T + i) * HD + h * AT_DK;
    const bool valid = i < len;
    for (int c16 = ch; c16 < AT_DK / 16; c16 += 2) {
      float ov[16];
      tmem_ld16(lane_base + (uint32_t)(c16 * 16), ov);  // warp-collective: rows past the tensor take part too
      if (i < T) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float4 o4 = valid ? make_float4(ov[e] * inv, ov[e + 1] * inv, ov[e + 2] * inv, ov[e + 3] * inv)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(orow + c16 * 16 + e) = o4;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// returns false when the shape is outside the fused kernel's range (the caller falls back to the CUDA-core kernels)
bool relpos_attention_umma(const float* q, const float* k, const float* v, const float* p, const float* bias_u,
                           const float* bias_v, const int64_t* lens, int B, int T, int H, int dk, int legacy, float* out,
                           int ld_qkv, cudaStream_t s) {
  static const bool disabled = [] { const char* e = getenv("PTTSPP_ATTN_TC"); return e && e[0] == '0'; }();
  if (disabled || dk != AT_DK || T > AT_MAXT || T < 1 || B > 65535 || H > 65535) return false;
  auto a16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
  if (!(a16(q) && a16(k) && a16(v) && a16(p) && a16(bias_u) && a16(bias_v) && a16(out) && ld_qkv % 4 == 0)) return false;
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  static int setup_dev[64] = {0};
  if (dev < 64 && !setup_dev[dev]) {
    PT_CUDA(cudaFuncSetAttribute(relpos_attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
    setup_dev[dev] = 1;
  } else if (dev >= 64) {
    PT_CUDA(cudaFuncSetAttribute(relpos_attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
  }
  const int Tp = legacy ? T : 2 * T - 1;
  dim3 grid(ceil_div(T, 128), H, B);
  relpos_attention_umma_kernel<<<grid, AT_THREADS, AT_SMEM, s>>>(q, k, v, p, bias_u, bias_v, lens, T, Tp, H, ld_qkv, legacy,
                                                                 1.f / sqrtf((float)dk), out);
  PT_LAUNCHED();
  return true;
}

}  // namespace pttspp
