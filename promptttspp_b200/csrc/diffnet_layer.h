// Fused DiffNet residual-layer stack (csrc/diffnet_layer.cu): host-side handle and the kernel's argument blocks.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.h"

namespace pttspp {

// per-layer constants in device memory (tensor maps of the packed split-fp16 weight planes, read by TMA from global)
struct alignas(64) DiffLayerConst {
  CUtensorMap wd_h, wd_l;  // dilated conv planes [3 * 512][256], gate/filter interleaved rows, box {64, 64}
  CUtensorMap wo_h, wo_l;  // output projection planes [512][256] (residual | skip rows)
  const float* bias_d;     // [512] interleaved order
  const float* bias_o;     // [512]
  float scale_d, scale_o;  // 2^-s of the power-of-two weight scales
  int dil;
};

struct DiffNetArgs {
  const DiffLayerConst* layers;
  int layer_begin, layer_end, n_layers_total;
  int B, T, n_mt, n_units;
  const float* cond;  // [layers][B][T][512] conditioner projections (+ bias), gate/filter interleaved
  int64_t cond_layer_stride;
  const float* step_emb;  // [layers][256] diffusion-step embedding of the current step
  __half* y_hi[2];        // ping-pong operand planes [B][T][256]: layer l reads y[l & 1], writes y[(l + 1) & 1]
  __half* y_lo[2];
  float* skip;            // [B][T][256] running skip sum
  __half* skip_hi;        // optional: the last layer writes the skip sum as operand planes instead
  __half* skip_lo;
  unsigned* done;         // [layers][units] completion counters
  unsigned done_target;
  float* dbg_z;           // tests: [B][T][256] gate output of the (single) layer run
  unsigned long long* dbg_prof;  // optional [clusters][16] cycle counters of the barrier waits (tools/bench_diffnet.py)
  int dbg_flags;          // timing experiments (PTTSPP_DIFFNET_DBG): 1 no epilogue global traffic, 2 one MMA per product,
                          // 4 no output projection, 256 half the weight traffic (lo planes not loaded: results invalid)
};

struct DiffLayerHost {
  const void *wd_hi, *wd_lo, *wo_hi, *wo_lo;  // device pointers
  const float *bias_d, *bias_o;
  float scale_d, scale_o;
  int dil;
};

struct DiffNetRun {
  int B, T, layer_begin, layer_end;
  const float* cond;
  const float* step_emb;
  void* y_hi[2];
  void* y_lo[2];
  float* skip;
  void* skip_hi;
  void* skip_lo;
  unsigned* done;  // flags_bytes(); zeroed by the caller before epoch 1
  unsigned epoch;  // 1, 2, ...: incremented per launch over the same `done` array
  float* dbg_z;
  unsigned long long* dbg_prof = nullptr;
};

struct DiffNetStack {
  void* d_layers = nullptr;
  int n_layers = 0;
  int setup_dev = -1, max_clusters = 0;
  CUtensorMap y_maps[4];
  const void* y_key[4] = {nullptr, nullptr, nullptr, nullptr};
  int y_B = 0, y_T = 0;
  ~DiffNetStack();
  void set_layers(const std::vector<DiffLayerHost>& layers);
  size_t flags_bytes(int B, int T) const;
  void run(const DiffNetRun& r, cudaStream_t s);
  static bool supported(int channels, int kernel, int max_dil) { return channels == 256 && kernel == 3 && max_dil <= 8; }
};

// ---- step boundary (csrc/diffnet_tail.cu): skip projection -> output projection -> DDPM update -> next input projection ----
struct alignas(64) DiffTailConst {
  CUtensorMap wsp_h, wsp_l;  // skip_projection planes [256][256], box {64, 64}
  CUtensorMap wop_h, wop_l;  // output_projection planes [mel][256] (rows >= mel zero-filled by TMA)
  CUtensorMap wip_h, wip_l;  // input_projection planes [256][mel padded to 128]
  const float *bias_sp, *bias_op, *bias_ip;
  float scale_sp, scale_op, scale_ip;  // accumulator scales: 2^-s of the weight scale (x 1/sqrt(layers) for the skip sum)
};

struct DiffTailArgs {
  const DiffTailConst* c;
  int B, T, n_mt, n_units, M;
  float* x;                // [B][T][M] x_t in, x_{t-1} out
  const float* z;          // [B][M][T] Gaussian noise of this step, or NULL (sigma = 0)
  float c_recip, c_recipm1, coef1, coef2, sigma;
  const float* step_next;  // [256] step embedding (layer 0) of the NEXT diffusion step; NULL: no input projection
  __half *y_hi, *y_lo;     // [B][T][256] operand planes of the next step's first layer
};

struct DiffTailHost {
  const void *wsp_hi, *wsp_lo, *wop_hi, *wop_lo, *wip_hi, *wip_lo;  // device pointers
  const float *bias_sp, *bias_op, *bias_ip;
  float scale_sp, scale_op, scale_ip;
  int mel, mel_pad;
};

struct DiffTailRun {
  int B, T;
  const void *skip_hi, *skip_lo;
  float* x;
  const float* z;
  float c_recip, c_recipm1, coef1, coef2, sigma;
  const float* step_next;
  void *y_hi, *y_lo;
};

struct DiffNetTail {
  void* d_const = nullptr;
  int mel = 0;
  int setup_dev = -1, max_clusters = 0;
  CUtensorMap s_maps[2];
  const void* s_key[2] = {nullptr, nullptr};
  int s_B = 0, s_T = 0;
  ~DiffNetTail();
  void set_weights(const DiffTailHost& h);
  void run(const DiffTailRun& r, cudaStream_t s);
};

}  // namespace pttspp
