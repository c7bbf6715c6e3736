// Shared helpers for the pttspp_b200 CUDA sources (sm_100a only).
#pragma once
#ifdef __CUDACC__
#include <cuda_fp16.h>
#endif
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pttspp_b200.h"

namespace pttspp {

// ---- error plumbing: C++ exceptions inside, int status + last_error at the C ABI ----
void set_last_error(const std::string& msg);
extern thread_local int64_t g_launch_count;

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define PT_CHECK(cond, ...)                                                          \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      char _buf[512];                                                                \
      snprintf(_buf, sizeof(_buf), __VA_ARGS__);                                     \
      throw ::pttspp::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + \
                            ": " + _buf);                                            \
    }                                                                                \
  } while (0)

#define PT_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw ::pttspp::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                            #expr + " -> " + cudaGetErrorString(_e));                      \
  } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define PT_LAUNCHED()              \
  do {                             \
    ++::pttspp::g_launch_count;    \
    PT_CUDA(cudaGetLastError());   \
  } while (0)

#define PT_API_BEGIN try {
#define PT_API_END                        \
  }                                       \
  catch (const std::exception& e) {       \
    ::pttspp::set_last_error(e.what());   \
    return 1;                             \
  }                                       \
  catch (...) {                           \
    ::pttspp::set_last_error("unknown");  \
    return 1;                             \
  }                                       \
  return 0;

// ---- optional per-launch timing (bench.py's roofline leg; off by default) -----------------
// When enabled, a ProfScope records a CUDA event pair around the launches issued inside it and
// accounts the ALGORITHMIC work of the call (flops / bytes as defined in DESIGN.md).
enum ProfTag { PROF_CONV_SIMT = 0, PROF_CONV_UMMA = 1, PROF_AA_SNAKE = 2, PROF_LAYERNORM = 3, PROF_ATTENTION = 4,
               PROF_OTHER = 5, PROF_DIFFNET = 6, PROF_NUM_TAGS = 7 };
struct ProfScope {
  int tag;
  cudaStream_t s;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(int tag, cudaStream_t s, double flops, double bytes);
  ~ProfScope();
};

#ifdef __CUDACC__
// fp32 -> fp16 with saturation to the largest finite half (one F2FP.SATFINITE, the cost of the plain conversion): the
// hi plane of a split-fp16 operand never becomes inf, so an activation beyond fp16's range (|v| >= 65520) degrades to a
// clipped value (exact up to 2 * 65504 through the lo plane) instead of inf - inf = NaN downstream.
__device__ __forceinline__ __half pt_f2h_sat(float v) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}
__device__ __forceinline__ __half2 pt_f2h2_sat(float a, float b) {  // .x = a, .y = b
  unsigned r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return *reinterpret_cast<__half2*>(&r);
}
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- internal op API (all on channels-last fp32 activations) --------------------------
void conv1d_cl(const pttspp_conv1d_desc& d, cudaStream_t s);
void pack_conv_weight(const float* v, const float* g, int Cout, int Cin, int K, float* packed, int w_ld,
                      int interleave_halves, cudaStream_t s);
void pack_conv_weight_split(const float* v, const float* g, int Cout, int Cin, int K, void* w_hi, void* w_lo,
                            int interleave_halves, float* scale_inv);
void split_f16_planes(const float* x, const float* add, int64_t n, int C, void* hi, void* lo, cudaStream_t s);
// [B][T][C] rows -> planes; rows >= len[b] (len optional) become zeros
void split_f16_rows(const float* x, int B, int T, int C, const int64_t* len, void* hi, void* lo, cudaStream_t s);
void pack_convtr_weight(const float* v, const float* g, int Cin, int Cout, int Kt, int stride, float* packed,
                        int w_ld, cudaStream_t s);
// tcgen05 path, one launch with two epilogues: columns [0, Cout(d1)) follow d1, the next Cout(d2) columns follow d2
void conv1d_umma_dual_cl(const pttspp_conv1d_desc& d1, const pttspp_conv1d_desc& d2, cudaStream_t s);
void layernorm_cl(const pttspp_layernorm_desc& d, cudaStream_t s);
// y (fp32) and/or y_hi/y_lo (split-fp16 planes, same indexing) receive the result
// symmetric_filters: the caller has checked f[k] == f[11-k] for both filters on the host (selects the channel-pair kernel)
void aa_snake_cl(const float* x, float* y, int B, int L, int C, const float* log_alpha, const float* up_f,
                 const float* down_f, cudaStream_t s, void* y_hi = nullptr, void* y_lo = nullptr,
                 int symmetric_filters = 0);
// Fused AA-Snake -> conv (conv1d_umma.cu): d.in = fp32 PRE-activation rows [B][T][C]; the activated operand planes are
// produced inside the conv kernel.  Needs exactly symmetric filters (as the channel-pair activation kernel).
bool aa_conv1d_supported(const pttspp_conv1d_desc& d);
void aa_conv1d_cl(const pttspp_conv1d_desc& d, const float* log_alpha, const float* up_f, const float* down_f,
                  cudaStream_t s);
void duration_quantize(const float* log_d, const int64_t* phone_len, int B, int Tx, int64_t* dur,
                       int64_t* frame_len, cudaStream_t s);
void length_regulate(const float* x, const int64_t* dur, int B, int Tx, int C, int Ty, float* out,
                     int32_t* idx_out, cudaStream_t s);
void relpos_attention(const float* q, const float* k, const float* v, const float* p, const float* bias_u,
                      const float* bias_v, const int64_t* lens, int B, int T, int H, int dk, int legacy,
                      float* scratch, float* out, int ld_qkv, cudaStream_t s);

// small element-wise / gather kernels of the acoustic model (acoustic_ops.cu)
void transpose_bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t s);
void transpose_btc_to_bct(const float* in, float* out, int B, int T, int C, const int64_t* len, float scale,
                          cudaStream_t s);
void transpose_btc_to_bct_affine(const float* in, float* out, int B, int T, int C, const int64_t* len, float scale,
                                 float shift, cudaStream_t s);
void embedding_cl(const int64_t* ids, const int64_t* len, const float* table, int B, int T, int C, int vocab,
                  float scale, float* out, cudaStream_t s);
void glu_dw_bn_swish_cl(const float* in /*[B][T][2C]*/, const int64_t* len, const float* dw_w /*[C][K]*/,
                        const float* dw_b, const float* bn_scale, const float* bn_shift, int B, int T, int C,
                        int K, float* out, cudaStream_t s);
void l2_normalize_rows(float* x, int rows, int C, cudaStream_t s);
void style_mdn_sample(const float* logpi, const float* logsigma, const float* mu /*[B][G*D]*/, const float* z,
                      int B, int G, int D, float noise_scale, int normalize, float* style, cudaStream_t s,
                      const float* comp_u = nullptr);
void add_row_broadcast(float* x /*[B][T][C]*/, const float* v /*[B][C]*/, int B, int T, int C, cudaStream_t s);
void mdn_duration_head(const float* h /*[B][T][C]*/, const float* w_pi, const float* b_pi, const float* w_ls,
                       const float* b_ls, const float* w_mu, const float* b_mu, int rows, int C, int G,
                       float* log_d, cudaStream_t s);
void pitch_head(const float* h /*[B][T][C]*/, const float* w /*[2][C]*/, const float* b, const int64_t* len,
                int B, int T, int C, float* log_cf0 /*[B][T]*/, float* vuv, cudaStream_t s);
void pitch_embed_add(float* x, const float* log_cf0, const float* w, const float* b, const int64_t* len, int B,
                     int T, int C, cudaStream_t s);
void ddpm_update(float* x /*[B][T][M]*/, const float* eps /*[B][T][M]*/, const float* z /*[B][M][T]*/, int B,
                 int T, int M, float c_recip, float c_recipm1, float coef1, float coef2, float sigma,
                 cudaStream_t s, void* xp_hi = nullptr, void* xp_lo = nullptr, int Mp = 0);
// [rows][C] fp32 -> split-fp16 planes [rows][Cp] (Cp >= C, columns >= C written as zeros)
void split_f16_pad(const float* x, int64_t rows, int C, int Cp, void* hi, void* lo, cudaStream_t s);


// ---- host-side tensor store + device buffers shared by the model-level handles -------------
struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const { return (int64_t)data.size(); }
};

struct TensorStore {
  std::map<std::string, HostTensor> t;
  void set(const char* name, const float* data, const int64_t* shape, int ndim, cudaStream_t s);
  bool has(const std::string& name) const { return t.count(name) != 0; }
  const HostTensor& get(const std::string& name) const;
  // get + shape check (numel)
  const HostTensor& get(const std::string& name, int64_t numel) const;
};

struct DeviceBuffers {
  std::vector<void*> ptrs;
  ~DeviceBuffers();
  void release();
  float* upload(const std::vector<float>& host);
  float* upload(const float* host, size_t n);
  void* upload_bytes(const void* host, size_t bytes);
};

struct PackedConv {
  float* w = nullptr;
  float* bias = nullptr;
  int Cin = 0, Cout = 0, K = 1, dil = 1, pad = 0, w_ld = 0;
  // optional split-fp16 copy for the tcgen05 path: [K][Cout][Cin] halves of w * 2^s, w_scale_inv = 2^-s
  void* w_hi = nullptr;
  void* w_lo = nullptr;
  float w_scale_inv = 0.f;
  float* bias_cols = nullptr;  // bias in packed column order (== bias)
};
// adds the split-fp16 planes of the same torch weight to an already loaded PackedConv
void attach_split_weights(const TensorStore& st, DeviceBuffers& dev, const std::string& prefix, PackedConv& c,
                          bool interleave_halves);

// Conv1d weights `prefix.weight` or weight-norm pair `prefix.weight_g/_v`, plus `prefix.bias`.
PackedConv load_conv1d(const TensorStore& st, DeviceBuffers& dev, const std::string& prefix, int Cout, int Cin, int K,
                       int dil, int pad, bool interleave_halves = false, bool has_bias = true);
// nn.Linear `prefix.weight` [Cout][Cin] (+ bias) as a K=1 conv.
PackedConv load_linear(const TensorStore& st, DeviceBuffers& dev, const std::string& prefix, int Cout, int Cin,
                       bool has_bias = true);
pttspp_conv1d_desc conv_desc(const PackedConv& c, const float* in, int B, int T, float* out);

}  // namespace pttspp
