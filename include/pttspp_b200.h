/*
 * pttspp_b200.h -- C ABI of the B200-native (sm_100a) PromptTTS++ inference hot path.
 *
 * The reference (line/promptttspp) has no FFI of its own: its "plugin" boundary is the
 * Python object protocol that Hydra `instantiate` + app.py / egs/proposed/bin/synthesize.py
 * rely on.  This header is the C boundary *beneath* that protocol:
 * every entry point takes plain pointers, sizes and a cudaStream_t (as void*), returns an
 * int status (0 = ok) and leaves a message for pttspp_last_error() otherwise.  The Python
 * shims in promptttspp_b200/ bind it with ctypes (INTEGRATION.md shows the stub).
 *
 * Activation layout ("cl" = channels-last): x[b][t][c], c contiguous, fp32.
 * Reference-facing tensors keep the reference layout ([B, C, T]) and are transposed inside.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef PTTSPP_B200_H_
#define PTTSPP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTTSPP_ABI_VERSION 1

typedef void* pttspp_stream_t; /* cudaStream_t */

/* ---- status / introspection ------------------------------------------------------- */
const char* pttspp_last_error(void);
int pttspp_abi_version(void);
/* 0 when the current CUDA device is compute capability 10.x (the only supported target). */
int pttspp_device_check(void);
/* Number of kernel launches issued by this library on the calling thread since the last
 * reset (bench.py's `gpu_launches`). */
int64_t pttspp_launch_count(void);
void pttspp_reset_launch_count(void);

/* Developer knobs (PTTSPP_UMMA_* environment variables: kernel-variant selection for op tests and A/B measurements) are
 * read once; this re-reads them. */
void pttspp_debug_reload_env(void);

/* Optional per-launch timing for bench.py's roofline leg (off by default, single-threaded use).
 * While enabled, every op-level call records a CUDA event pair on its stream and accounts its
 * ALGORITHMIC work; the report sums them per kernel family:
 *   tag 0 conv1d CUDA-core, 1 conv1d tcgen05, 2 anti-aliased Snake, 3 LayerNorm, 4 attention, 5 other,
 *   6 the fused DiffNet layer-stack kernel (one launch = all residual layers of one diffusion step).
 * ms / flops / bytes / calls: arrays of >= 7 entries.  prof_enable() also clears the counters. */
void pttspp_prof_enable(int on);
int pttspp_prof_report(double* ms, double* flops, double* bytes, int64_t* calls, int ntags);

/* ---- op level -------------------------------------------------------------------- */

enum {
  PTTSPP_ACT_NONE = 0,
  PTTSPP_ACT_RELU = 1,
  PTTSPP_ACT_GELU = 2,  /* exact erf GELU (torch.nn.GELU(), frame_prior.py:64) */
  PTTSPP_ACT_SWISH = 3, /* x*sigmoid(x) (esp/conformer/swish.py:18) */
  PTTSPP_ACT_GATE = 4,  /* sigmoid(v[2j])*tanh(v[2j+1]) -> Cout/2 outputs (denoiser.py:76-77,
                           weights packed gate/filter interleaved) */
  PTTSPP_ACT_TANH = 5
};

/* Conv1d / Linear / polyphase ConvTranspose1d on channels-last activations with a fused
 * prologue (input length mask, per-channel input add) and epilogue:
 *     v   = act(acc_scale * conv(in)[row, co] + bias[co] + addend[row, co])
 *     out = (res_scale * res[row, co] + alpha * mask(row) * v + beta * out[row, co]) / out_div
 * for output index m in [m_begin, m_begin + M): row = m * out_mul + out_off, the input row of
 * tap k is m * in_stride + k * dil - pad; rows outside [0, T_in) (or >= in_len[b]) read as 0,
 * rows outside [0, T_out) are not written.
 * Replaces torch.nn.Conv1d / Linear / ConvTranspose1d call sites of the hot path, e.g.
 * promptttspp/modules/denoiser.py:69-83, esp/transformer/multi_layer_conv.py:65-67,
 * promptttspp/vocoders/bigvgan.py:42-47,90-102. */
typedef struct {
  const float* in;
  int64_t in_bs; /* batch stride, elements */
  int32_t in_ld; /* row stride, elements (multiple of 4) */
  int32_t T_in;
  int32_t Cin; /* multiple of 16 */
  const float* w; /* packed [K][Cin][w_ld] (pttspp_pack_conv_weight) */
  int32_t w_ld;   /* multiple of 4, >= Cout */
  const float* bias; /* [Cout] or NULL */
  int32_t K, dil, pad, in_stride;
  float* out;
  int64_t out_bs;
  int32_t out_ld;
  int32_t T_out;
  int32_t Cout; /* pre-activation columns */
  int32_t m_begin, M;
  int32_t out_mul, out_off;
  const int64_t* in_len;  /* optional [B] */
  const int64_t* out_len; /* optional [B]: mask(row) = row < out_len[b] */
  const float* in_add;    /* optional [Cin], added to rows inside [0, min(T_in, in_len)) */
  const float* addend;    /* optional, indexed like out but with Cout columns */
  int64_t addend_bs;
  int32_t addend_ld;
  int32_t act;
  float acc_scale;
  const float* res; /* optional, indexed like out */
  int64_t res_bs;
  int32_t res_ld;
  float res_scale;
  float alpha, beta;
  float out_div; /* 0 = no division */
  int32_t B;
  int32_t impl; /* 0 = auto, 1 = force SIMT fp32, 2 = force tcgen05 split-fp16, 3 = tcgen05 split-fp16 with chunked
                 * near-fp32 accumulation (<= 16 tensor-core accumulations per accumulator, chunk sums added in
                 * round-to-nearest fp32 registers): the text side, whose outputs decide integer durations;
                 * 4 = tcgen05 split-fp16 without chunking: contractions of up to 256 accumulations per accumulator and
                 * single-N-tile convs may take the CTA-pair kernel (the vocoder, whose bar is 1e-4 RMS on the waveform) */
  /* ---- split-fp16 operands (tcgen05 path) -------------------------------------------------
   * An fp32 value v travels as two fp16 planes hi = fp16(v), lo = fp16(v - hi); the tensor cores
   * compute hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (~2^-22 relative, fp32 class).
   * in_hi/in_lo: [B][T_in][Cin] halves (row stride Cin, batch stride T_in*Cin), replace `in`;
   * w_hi/w_lo: packed [K][Cout][Cin] halves of w * 2^s (pttspp_pack_conv_weight_split), w_scale_inv = 2^-s.
   * out_hi/out_lo (optional, any path): planes of (out + out_plane_add[c]) for the next conv;
   * `out` may be NULL then. */
  const void* in_hi;
  const void* in_lo;
  const void* w_hi;
  const void* w_lo;
  float w_scale_inv;
  void* out_hi;
  void* out_lo;
  const float* out_plane_add; /* optional [output columns] */
  int64_t out_plane_bs;
  int32_t out_plane_ld;
  /* Residual taken from split-fp16 planes instead of `res` (which must then be NULL):
   *   r[row, c] = float(res_hi[row, c]) + float(res_lo[row, c]) - res_plane_sub[c]
   * DiffNet's residual stream h travels only as the operand planes y = h + step_emb[l] of the next dilated conv; the
   * 1x1 output projection recovers h from them (denoiser.py:79-83) and no fp32 copy of h is stored.  Supported by the
   * CTA-pair tcgen05 kernel (which such a launch always takes). */
  const void* res_hi;
  const void* res_lo;
  const float* res_plane_sub; /* optional [output columns] */
  int64_t res_plane_bs;
  int32_t res_plane_ld;
} pttspp_conv1d_desc;

int pttspp_conv1d_cl(const pttspp_conv1d_desc* d, pttspp_stream_t stream);
/* tcgen05 path only: one contraction, two epilogues.  Output columns [0, d1->Cout) follow d1, the next d2->Cout columns
 * follow d2 (its bias/res/out pointers are relative to its own first column).  Both must share the input planes and
 * geometry, K == 1, d1->Cout % 128 == 0 and d2's weight planes must directly follow d1's.  Used for the residual |
 * skip halves of DiffNet's output projection (denoiser.py:79-83). */
int pttspp_conv1d_dual_cl(const pttspp_conv1d_desc* d1, const pttspp_conv1d_desc* d2, pttspp_stream_t stream);

/* Weight packing runs once at load time on the HOST (v, g, packed are host pointers).
 * Repack a torch Conv1d weight [Cout][Cin][K] into [K][Cin][w_ld];
 * if g != NULL the weight-norm reparametrisation w = g * v / ||v|| (norm over dims 1,2;
 * torch.nn.utils.weight_norm dim=0, bigvgan.py:24-36) is folded in.  `interleave_halves`
 * packs output channel c of the first half to column 2c and of the second half to 2c+1. */
int pttspp_pack_conv_weight(const float* v, const float* g, int Cout, int Cin, int K, float* packed,
                            int w_ld, int interleave_halves, pttspp_stream_t stream);
/* Split-fp16 packing for the tcgen05 path: w_hi/w_lo [K][Cout][Cin] halves of w * 2^s with the largest
 * power of two that keeps max|w| * 2^s <= 16384; returns 2^-s in *scale_inv.  Host pointers. */
int pttspp_pack_conv_weight_split(const float* v, const float* g, int Cout, int Cin, int K, void* w_hi, void* w_lo,
                                  int interleave_halves, float* scale_inv);
/* x[n] (fp32, device) -> hi[n], lo[n] (fp16, device) planes of (x + add[i % C]) (add optional). */
int pttspp_split_f16(const float* x, const float* add, int64_t n, int C, void* hi, void* lo, pttspp_stream_t stream);

/* Hardware probe (tests only): D[128][128] = A[row_off:row_off+128][0:64] . B[0:128][0:64]^T on tcgen05 with the A tile
 * loaded once (144 rows, fp16) and the shared-memory descriptor start advanced by row_off rows; mode 1 sets the
 * descriptor's base_offset field to (start >> 7) & 7. */
int pttspp_umma_probe(const void* a_half, int rows, const void* b_half, int row_off, int mode, float* out,
                      pttspp_stream_t stream);

/* Repack a ConvTranspose1d weight [Cin][Cout][Kt] (+ optional weight-norm g[Cin]) into `stride`
 * polyphase 2-D conv weights [stride][Kt/stride][Cin][w_ld] (bigvgan.py:90-102). */
int pttspp_pack_convtr_weight(const float* v, const float* g, int Cin, int Cout, int Kt, int stride,
                              float* packed, int w_ld, pttspp_stream_t stream);

/* LayerNorm over the channel dim of x[b][t][c]:
 *   x = in_scale * (in * mask_in) + in2 + row_add[t];  y = LN(x) * gamma + beta;  out = y * mask_out
 * Replaces esp/transformer/layer_norm.py:21 (eps 1e-12), layers/norm.py:26-32, frame_prior.py:31-34
 * and the `x*16 + pe` of modules/embedding.py:90-92. */
typedef struct {
  const float* in;
  const float* in2;     /* optional, same indexing as in */
  const float* row_add; /* optional [T][C] */
  const float* gamma;
  const float* beta;
  float* out;
  int64_t bs; /* batch stride (elements) of in/in2/out */
  int32_t ld;
  int32_t B, T, C; /* C multiple of 4, <= 1024 */
  float eps, in_scale;
  const int64_t* in_len;
  const int64_t* out_len;
} pttspp_layernorm_desc;
int pttspp_layernorm_cl(const pttspp_layernorm_desc* d, pttspp_stream_t stream);

/* Fused anti-aliased Snake: 2x polyphase up-sample (replicate pad) -> x + sin^2(a x)/(a + 1e-9)
 * -> 2x low-pass down-sample, never materialising the 2x signal.  x, y: [B][L][C] channels-last.
 * Replaces promptttspp/layers/activations.py:22-33 (UpSample1d :74-96, Snake :36-44,
 * DownSample1d :99-138).  up_filter / down_filter: 12 taps each (device pointers, the
 * checkpoint's `up.filter` / `down.lowpass.filter` buffers); log_alpha: [C]. */
int pttspp_aa_snake_cl(const float* x, float* y, int B, int L, int C, const float* log_alpha,
                       const float* up_filter, const float* down_filter, pttspp_stream_t stream);
/* The same activation through the channel-pair kernel (two adjacent channels per thread in packed fp32x2 registers, a
 * rolling strip without halo recomputation): what the BigVGAN handle uses.  PRECONDITION: both filters are exactly
 * symmetric, f[k] == f[11-k] (true for the reference's Kaiser sinc, layers/activations.py:36-71); C even.  Every
 * channel's result is then bit-identical to pttspp_aa_snake_cl's. */
int pttspp_aa_snake_pair_cl(const float* x, float* y, int B, int L, int C, const float* log_alpha,
                            const float* up_filter, const float* down_filter, pttspp_stream_t stream);

/* AA-Snake fused into the conv that consumes it (BigVGAN's AMP layer, vocoders/bigvgan.py:42-47: `conv(act(x))`):
 * d->in is the fp32 PRE-activation tensor [B][T][C]; activation-producer warps inside the tcgen05 conv kernel run the
 * anti-aliased Snake (same arithmetic and order as pttspp_aa_snake_pair_cl) and write the split-fp16 operand planes
 * straight into shared memory, so the activated tensor never exists in HBM.  Result is bit-identical to
 * pttspp_aa_snake_pair_cl followed by pttspp_conv1d_cl on the planes.  Supported: Cin == Cout == 32 (K <= 11) or 64
 * (K <= 7), stride 1, split-fp16 weights (w_hi / w_lo / w_scale_inv), 32-byte aligned out / res, symmetric filters. */
int pttspp_aa_conv1d_cl(const pttspp_conv1d_desc* d, const float* log_alpha, const float* up_filter,
                        const float* down_filter, pttspp_stream_t stream);

/* Duration quantisation + length regulator.
 *   dur[b][i] = (i < len[b]) ? max(rint(exp(log_d[b][i])), 1) : 0    (variance_adaptor.py:179-181)
 *   frame_len[b] = sum_i dur[b][i]                                   (variance_adaptor.py:183)
 * then out[b][t][:] = x[b][idx(b,t)][:] for t < frame_len[b] else 0, idx = searchsorted(cumsum(dur), t,
 * right) -- the gather that `x @ generate_path(...)` (utils/model.py:37-47,
 * variance_adaptor.py:185-187) computes with a dense one-hot matmul. */
int pttspp_duration_quantize(const float* log_d, const int64_t* phone_len, int B, int Tx, int64_t* dur,
                             int64_t* frame_len, pttspp_stream_t stream);
int pttspp_length_regulate(const float* x, const int64_t* dur, int B, int Tx, int C, int Ty, float* out,
                           int32_t* idx_out /* optional [B][Ty], -1 on padding */, pttspp_stream_t stream);

/* Zero-phase IIR filter of rows x[rows][T] -> y (scratch: rows*T floats): torchaudio.functional.filtfilt(x, a, b,
 * clamp=False), i.e. lfilter forward, flip, lfilter, flip, zero initial state, coefficients normalised by a[0]; b/a:
 * ntaps (<= 8) device floats.  Replaces the torch branch of promptttspp/utils/model.py:164-196 (the 5th-order
 * Butterworth low-pass app.py:77 applies to log-f0 between the acoustic model and the vocoder). */
int pttspp_iir_filtfilt(const float* x, float* y, float* scratch, int rows, int T, const float* b_coeffs,
                        const float* a_coeffs, int ntaps, pttspp_stream_t stream);
/* The same filter over a zero-padded ragged batch: row r holds row_len[r] valid samples and is filtered exactly like a
 * per-utterance call of promptttspp/utils/model.py:164-196 on x[r, :row_len[r]] (the backward pass starts at the row's own
 * last sample with zero state); rows with row_len <= min_len pass through unchanged (the reference's short-input rule,
 * utils/model.py:186-187) and the padding is copied.  This is what a batched caller of app.py:76-77 needs. */
int pttspp_iir_filtfilt_ragged(const float* x, float* y, float* scratch, int rows, int T, const int64_t* row_len,
                               int min_len, const float* b_coeffs, const float* a_coeffs, int ntaps,
                               pttspp_stream_t stream);

/* Log-mel spectrogram front end of the reference-mel path: promptttspp/transforms/mel.py:18-34 (torchaudio
 * MelSpectrogram as configured by conf/transforms/mel.yaml; callers app.py:93-100, egs/proposed/bin/synthesize.py:172-175).
 * wav [B][L] -> frames = 1 + L / hop columns; frame t covers samples t*hop - n_fft/2 .. + n_fft (reflect padding at both
 * ends, torch.stft(center=True, pad_mode="reflect")), multiplied by window[n_fft] (the analysis window already
 * zero-padded to n_fft), |DFT|^power (power 1 or 2) over n_fft/2 + 1 bins -> spec_out [B][n_fft/2+1][frames] (optional);
 * mel_out [B][n_mels][frames] = log(max(fb^T spec, log_floor)) with fb [n_fft/2+1][n_mels] (optional). */
int pttspp_mel_spectrogram(const float* wav, int B, int L, int n_fft, int hop, const float* window, const float* fb,
                           int n_mels, int power, float log_floor, float* spec_out, float* mel_out,
                           pttspp_stream_t stream);
/* spec_to_mel alone (transforms/mel.py:23-26): spec [B][n_freq][frames] -> mel_out [B][n_mels][frames]. */
int pttspp_mel_from_spec(const float* spec, int B, int n_freq, int frames, const float* fb, int n_mels, float log_floor,
                         float* mel_out, pttspp_stream_t stream);

/* Reference-mel style path (promptttspp/modules/reference_encoder.py:95-124, style_encoder.py:82-171).
 * conv2d_bn_relu: one [Conv2d KxK stride pad (no bias) -> BatchNorm2d(eval, folded to scale/shift) -> ReLU] block on
 *   NCHW tensors x [B][Cin][H][W] -> out [B][Cout][Ho][Wo] (reference_encoder.py:66-81).
 * gru_last_state: torch.nn.GRU(I, Hn, 1, batch_first) over x [B][T][I]; out [B][Hn] = hidden state after the last valid
 *   step t < lens[b] (pack_padded_sequence + `_, ref_embs = self.gru(hs)`, :112-121); lens NULL = T; gates (r, z, n).
 * style_token_attention: q = Linear(ref [B][R]); k, v = Linear(tanh(gst_embs [Tk][Dk])); `heads` heads over Fdim
 *   features, scores / sqrt(Fdim), softmax over tokens, linear_out -> out [B][Fdim] (style_encoder.py:123-171). */
int pttspp_conv2d_bn_relu(const float* x, const float* w, const float* bn_scale, const float* bn_shift, float* out, int B,
                          int Cin, int H, int W, int Cout, int K, int stride, int pad, pttspp_stream_t stream);
int pttspp_gru_last_state(const float* x, const int64_t* lens, int B, int T, int I, int Hn, const float* w_ih,
                          const float* w_hh, const float* b_ih, const float* b_hh, float* out, pttspp_stream_t stream);
int pttspp_style_token_attention(const float* ref, int B, int R, const float* gst_embs, int Tk, int Dk, int heads, int Fdim,
                                 const float* wq, const float* bq, const float* wk, const float* bk, const float* wv,
                                 const float* bv, const float* wo, const float* bo, float* out, pttspp_stream_t stream);

/* BERT building blocks for the prompt encoder's sentence embedding (promptttspp/modules/prompt_encoder.py:22-38 calls HF
 * transformers' BertModel; third-party).  bert_embed: out[b][t] = word_emb[ids[b][t]] + type_emb0 + pos_emb[t]
 * (BertEmbeddings before its LayerNorm).  mha_masked: BertSelfAttention on a fused q|k|v buffer qkv [B][T][3*heads*dk]:
 * softmax(q k^T / sqrt(dk) + (1 - key_mask) * finfo.min) v -> out [B][T][heads*dk]; key_mask [B][T] int64 (1 = token) or
 * NULL.  Linear / GELU / LayerNorm layers use pttspp_conv1d_cl and pttspp_layernorm_cl. */
int pttspp_bert_embed(const int64_t* ids, int B, int T, const float* word_emb, int vocab, const float* pos_emb,
                      const float* type_emb0, int hidden, float* out, pttspp_stream_t stream);
int pttspp_mha_masked(const float* qkv, const int64_t* key_mask, int B, int T, int heads, int dk, float* out,
                      pttspp_stream_t stream);

/* Fused DiffNet residual-layer stack on tcgen05 (one persistent cta_group::2 kernel per launch, layers chained inside the
 * kernel by completion flags).  Per layer l, on split-fp16 operand planes y = h + step_emb[l]:
 *     g|f = dilated_conv_l(y) + bias_d + cond_l;  z = sigmoid(g) * tanh(f);  r|s = W_o z + bias_o;
 *     y' = (h + r) / sqrt(2) + step_emb[l+1];  skip += s
 * z stays in tensor memory (operand of the second contraction), h only travels as the planes.
 * Replaces promptttspp/modules/denoiser.py:69-83 (ResidualBlock.forward) for residual_channels = 256, kernel 3,
 * dilation <= 8.  All pointers are device pointers, 32-byte aligned.
 *   wd_hi/wd_lo: [3][512][256] halves, gate/filter rows interleaved (pttspp_pack_conv_weight_split, interleave_halves=1);
 *   wo_hi/wo_lo: [512][256] halves (residual rows, then skip rows); scale_* = the packers' scale_inv;
 *   bias_d: [512] in the interleaved order, bias_o: [512]. */
typedef struct {
  const void* wd_hi;
  const void* wd_lo;
  const void* wo_hi;
  const void* wo_lo;
  const float* bias_d;
  const float* bias_o;
  float scale_d, scale_o;
  int32_t dil;
} pttspp_diffnet_layer;
typedef struct {
  int32_t B, T;                 /* utterances, frames per utterance (all rows are live: no masks, diffusion.py:199) */
  int32_t layer_begin, layer_end; /* layers [begin, end) of the stack run by this launch */
  const float* cond;            /* [layers][B][T][512] conditioner projections incl. bias, gate/filter interleaved */
  const float* step_emb;        /* [layers + 1][256]: diffusion-step embedding per layer (row `layers` is never read
                                   by the last layer, which has no residual output) */
  void* y_hi[2];                /* ping-pong planes [B][T][256] halves: layer l reads y[l & 1], writes y[(l + 1) & 1] */
  void* y_lo[2];
  float* skip;                  /* [B][T][256] running skip sum (layer 0 overwrites) */
  void* skip_hi;                /* optional: the last layer writes the skip sum as operand planes instead of `skip` */
  void* skip_lo;
  uint32_t* done;               /* pttspp_diffnet_flags_bytes() bytes, zeroed by the caller before epoch 1 */
  uint32_t epoch;               /* 1, 2, ...: one more per launch over the same `done` array */
  float* dbg_z;                 /* tests: [B][T][256] gate output z of the last layer run, or NULL */
  uint64_t* dbg_prof;           /* profiling: [74][16] cycle counters of the kernel's barrier waits, or NULL */
} pttspp_diffnet_run_desc;
typedef struct pttspp_diffnet pttspp_diffnet_t;
int pttspp_diffnet_create(const pttspp_diffnet_layer* layers, int n_layers, pttspp_diffnet_t** out);
void pttspp_diffnet_destroy(pttspp_diffnet_t* h);
size_t pttspp_diffnet_flags_bytes(const pttspp_diffnet_t* h, int B, int T);
int pttspp_diffnet_run(pttspp_diffnet_t* h, const pttspp_diffnet_run_desc* r, pttspp_stream_t stream);

/* Relative-position multi-head self-attention (Transformer-XL style), both ESPnet variants.
 *   scores = ((q+u) k^T + rel_shift((q+v) p^T)) / sqrt(d_k); masked softmax; . v
 * q,k,v,out: [B][T][H*d_k]; p: [Tp][H*d_k] with Tp = T (legacy) or 2T-1 (new);
 * bias_u, bias_v: [H][d_k]; lens: [B].  Replaces esp/transformer/attention.py:164-206 (legacy,
 * incl. the wrapped upper triangle of its rel_shift :142-162), :262-305 (new) and :63-93.
 * scratch: B*H*T*Tp floats. */
int pttspp_relpos_attention(const float* q, const float* k, const float* v, const float* p,
                            const float* bias_u, const float* bias_v, const int64_t* lens, int B, int T,
                            int H, int dk, int legacy, float* scratch, float* out, pttspp_stream_t stream);

/* ---- model level: BigVGAN ---------------------------------------------------------- */

typedef struct pttspp_bigvgan pttspp_bigvgan_t;
typedef struct {
  int32_t in_channel;               /* 80 */
  int32_t upsample_initial_channel; /* 512 */
  int32_t num_upsamples;            /* <= 8 */
  int32_t upsample_rates[8];
  int32_t upsample_kernel_sizes[8];
  int32_t num_kernels; /* <= 8 */
  int32_t resblock_kernel_sizes[8];
  int32_t num_dilations; /* layers per AMPBlock, <= 8 */
  int32_t resblock_dilations[8][8];
} pttspp_bigvgan_config;

/* promptttspp/vocoders/bigvgan.py:71-118 (constructor), conf/vocoder/bigvgan.yaml. */
int pttspp_bigvgan_create(const pttspp_bigvgan_config* cfg, pttspp_bigvgan_t** out);
void pttspp_bigvgan_destroy(pttspp_bigvgan_t* h);
/* One call per state_dict entry (reference key names: conv_pre.weight_g, mrfs.0.1.layers.2.act1.act.alpha,
 * ...; `weight` instead of weight_g/weight_v after remove_weight_norm_).  `data` may be a host or a
 * device pointer; the library keeps its own copy.  Replaces load_state_dict (app.py:35-37). */
int pttspp_bigvgan_set_tensor(pttspp_bigvgan_t* h, const char* name, const float* data, const int64_t* shape,
                              int ndim, pttspp_stream_t stream);
/* Folds weight-norm, repacks all weights.  Fails naming the first missing key. */
int pttspp_bigvgan_finalize(pttspp_bigvgan_t* h, pttspp_stream_t stream);
size_t pttspp_bigvgan_workspace_bytes(const pttspp_bigvgan_t* h, int B, int T);
/* mel: [B][in_channel][T] (reference layout), wav: [B][1][T*prod(rates)].
 * Replaces BigVGAN.forward (bigvgan.py:120-131). */
int pttspp_bigvgan_forward(pttspp_bigvgan_t* h, const float* mel, int B, int T, float* wav, void* workspace,
                           size_t workspace_bytes, pttspp_stream_t stream);

/* F0-aware generator (promptttspp/vocoders/bigvgan_f0.py:98-115): the handle must have received the
 * `noise_convs.{i}.{weight,bias}` tensors; har_source: [B][T*prod(rates)] from pttspp_nsf_source.  Per stage
 * x = up(x) + noise_conv(har_source) (:104-106), everything else as pttspp_bigvgan_forward. */
int pttspp_bigvgan_forward_f0(pttspp_bigvgan_t* h, const float* mel, const float* har_source, int B, int T, float* wav,
                              void* workspace, size_t workspace_bytes, pttspp_stream_t stream);
/* Harmonic-plus-noise source: nn.Upsample(scale_factor=hop) of f0 [B][T] (bigvgan_f0.py:99), SineGen
 * (promptttspp/vocoders/nsf.py:55-85 phase accumulation with torch's double-accumulated cumsum, :116-148 voiced /
 * unvoiced mix) and SourceModuleHnNSF's Linear(H -> 1) + tanh (:193-206).  The random draws of the reference are
 * INPUTS (device): rand_ini [B][H] (torch.rand, column 0 zeroed, nsf.py:64-67) and noise [B][L][H] (randn_like, :143);
 * H = harmonic_num + 1 <= 16, L = T*hop.  har_source: [B][L]. */
size_t pttspp_nsf_source_workspace_bytes(int B, int T, int hop, int harmonic_num);
int pttspp_nsf_source(const float* f0, int B, int T, int hop, float sampling_rate, int harmonic_num, float sine_amp,
                      float noise_std, float voiced_threshold, const float* rand_ini, const float* noise,
                      const float* lin_w, const float* lin_b, float* har_source, void* workspace, size_t workspace_bytes,
                      pttspp_stream_t stream);

/* ---- model level: acoustic model (PromptTTSMDNDurCFG inference) --------------------- */

typedef struct pttspp_acoustic pttspp_acoustic_t;
typedef struct {
  int32_t num_vocab;  /* 90 */
  int32_t channels;   /* 256 */
  int32_t emb_do_scale;
  /* Conformer encoder (conf/model/prompttts_mdn_v2_wo_erg_final.yaml:13-30) */
  int32_t enc_heads, enc_linear_units, enc_blocks, enc_ff_kernel, enc_cnn_kernel;
  int32_t rel_pos_legacy; /* 1 = legacy, 0 = new */
  /* variance adaptor (:32-64) */
  int32_t dur_layers, dur_kernel, dur_gaussians;
  int32_t pitch_layers, pitch_kernel;
  int32_t fp_layers, fp_kernel;
  /* prompt adaptor + style MDN (:79-91) */
  int32_t prompt_in, prompt_mid, style_gaussians;
  int32_t norm_style_emb;
  /* diffusion decoder (:93-105) */
  int32_t mel_dim, K_step, diff_layers, diff_channels, diff_kernel, diff_dilation_cycle;
  float diff_scale; /* SinusoidalPosEmb scale */
  float norm_scale; /* <= 0: use a_min/a_max */
  float a_min, a_max;
} pttspp_acoustic_config;

int pttspp_acoustic_create(const pttspp_acoustic_config* cfg, pttspp_acoustic_t** out);
void pttspp_acoustic_destroy(pttspp_acoustic_t* h);
int pttspp_acoustic_set_tensor(pttspp_acoustic_t* h, const char* name, const float* data,
                               const int64_t* shape, int ndim, pttspp_stream_t stream);
int pttspp_acoustic_finalize(pttspp_acoustic_t* h, pttspp_stream_t stream);

size_t pttspp_acoustic_encode_workspace_bytes(const pttspp_acoustic_t* h, int B, int Tx);
/* Text side of PromptTTSMDNDurCFG.infer_batch (models/prompttts_mdn_v2_final/model.py:261-303 and
 * variance_adaptor.py:178-183): embedding -> Conformer -> prompt adaptor + style MDN sample ->
 * x + style -> MDN duration predictor -> integer durations.
 *   phoneme [B][Tx] int64, phone_len [B] int64, pos_emb [Tp][C] (Tp = Tx legacy / 2Tx-1 new),
 *   cls_emb [B][prompt_in] (BERT CLS), z_style [B][C] (the randn_like draw of model.py:191),
 * outputs: enc_state [B][Tx][C] (x + style), dur [B][Tx] int64, frame_len [B] int64,
 *   log_dur [B][Tx] (optional), style_emb [B][C] (optional). */
int pttspp_acoustic_encode(pttspp_acoustic_t* h, const int64_t* phoneme, const int64_t* phone_len, int B,
                           int Tx, const float* pos_emb, int Tp, const float* cls_emb, const float* z_style,
                           float noise_scale, int use_max, float* enc_state, int64_t* dur,
                           int64_t* frame_len, float* log_dur, float* style_emb, void* workspace,
                           size_t workspace_bytes, pttspp_stream_t stream);

/* use_max = False (model.py:185-196 with mdn.py:226-257): the style MDN component of every dimension is a
 * Categorical(probs = exp(log_pi)) draw instead of the arg-max.  The draw is an input: comp_u [B][C] uniforms in [0, 1),
 * component = inverse CDF of the normalised probabilities. */
int pttspp_acoustic_encode_sampled(pttspp_acoustic_t* h, const int64_t* phoneme, const int64_t* phone_len, int B, int Tx,
                                   const float* pos_emb, int Tp, const float* cls_emb, const float* z_style,
                                   const float* comp_u, float noise_scale, float* enc_state, int64_t* dur,
                                   int64_t* frame_len, float* log_dur, float* style_emb, void* workspace,
                                   size_t workspace_bytes, pttspp_stream_t stream);
/* The same text side with the style vector given instead of derived from a prompt: style_in [B][C] is the output of
 * the reference-mel style encoder (model.py:232-235 / :297-301); it is L2-normalised here when norm_style_emb is set. */
int pttspp_acoustic_encode_ref(pttspp_acoustic_t* h, const int64_t* phoneme, const int64_t* phone_len, int B, int Tx,
                               const float* pos_emb, int Tp, const float* style_in, float* enc_state, int64_t* dur,
                               int64_t* frame_len, float* log_dur, void* workspace, size_t workspace_bytes,
                               pttspp_stream_t stream);

size_t pttspp_acoustic_decode_workspace_bytes(const pttspp_acoustic_t* h, int B, int Tx, int Ty);
/* Frame side (variance_adaptor.py:185-206, model.py:311-325, diffusion.py:320-356): length regulator ->
 * frame prior -> pitch predictor -> + pitch embedding -> K_step DDPM ancestral sampling with the DiffNet
 * denoiser -> mel.
 *   pe_abs [Ty][C] sinusoidal table (modules/embedding.py:57-78), x_T [B][mel][Ty] and
 *   z [K_step][B][mel][Ty] the Gaussian draws of diffusion.py:332 / :218 (z[0] is used at step K-1),
 * outputs (reference layouts): mel [B][mel][Ty], log_cf0 [B][1][Ty], vuv [B][1][Ty];
 *   cond_out [B][Ty][C] optional (decoder conditioning, for tests). */
int pttspp_acoustic_decode(pttspp_acoustic_t* h, const float* enc_state, const int64_t* dur,
                           const int64_t* frame_len, int B, int Tx, int Ty, const float* pe_abs,
                           const float* x_T, const float* z, float* mel, float* log_cf0, float* vuv,
                           float* cond_out, void* workspace, size_t workspace_bytes, pttspp_stream_t stream);

/* The same with the K_step noise tensors drawn INSIDE the loop, one [B][mel][Ty] buffer at a time, from torch's CUDA
 * Philox stream: (rng_seed, rng_offset) = the generator's (initial_seed(), get_offset()) after x_T was drawn; every
 * step's draw is bit-identical to the torch.randn(shape) the reference makes at diffusion.py:218, and *rng_offset_out is
 * the offset the caller must set the generator to afterwards.  Removes the [K_step][B][mel][Ty] tensor (1.3 GB at
 * batch 16 x 2.6 k frames). */
int pttspp_acoustic_decode_rng(pttspp_acoustic_t* h, const float* enc_state, const int64_t* dur,
                               const int64_t* frame_len, int B, int Tx, int Ty, const float* pe_abs,
                               const float* x_T, uint64_t rng_seed, uint64_t rng_offset, uint64_t* rng_offset_out,
                               float* mel, float* log_cf0, float* vuv, float* cond_out, void* workspace,
                               size_t workspace_bytes, pttspp_stream_t stream);
/* One torch-compatible standard-normal draw: out[numel] = what `torch.empty(numel, device="cuda").normal_()` writes when
 * the CUDA generator holds (seed, offset); *offset_advance = how far that call moves the generator's offset
 * (ATen/native/cuda/DistributionTemplates.h: grid-stride Philox4_32_10 + curand_normal4). */
int pttspp_philox_normal(float* out, int64_t numel, uint64_t seed, uint64_t offset, uint64_t* offset_advance,
                         pttspp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PTTSPP_B200_H_ */
