// Reference-mel style path: the three op-level kernels behind
// promptttspp_b200.modules.style_encoder.StyleEncoder.
//   ReferenceEncoder (promptttspp/modules/reference_encoder.py:95-124): 6 x [Conv2d k3 s2 (no bias) -> BatchNorm2d(eval)
//   -> ReLU] over the mel "image" [B, 1, L, 80], then a GRU over the sub-sampled time axis whose LAST VALID hidden state
//   is the reference embedding (pack_padded_sequence semantics);
//   StyleTokenLayer (promptttspp/modules/style_encoder.py:82-171): one query against tanh(gst_embs) with 4 heads.
// The whole path is a few MFLOP per utterance and runs once per call: plain CUDA-core kernels, no tensor cores.
#include "common.h"

namespace pttspp {
namespace {

// out[b][co][ho][wo] = relu(scale[co] * sum_{ci,kh,kw} x[b][ci][ho*s-p+kh][wo*s-p+kw] * w[co][ci][kh][kw] + shift[co])
__global__ void __launch_bounds__(256) conv2d_bn_relu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             float* __restrict__ out, int B, int Cin, int H, int W, int Cout,
                                                             int Ho, int Wo, int K, int stride, int pad) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Cout * Ho * Wo;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo);
  const int ho = (int)((idx / Wo) % Ho);
  const int co = (int)((idx / ((int64_t)Wo * Ho)) % Cout);
  const int b = (int)(idx / ((int64_t)Wo * Ho * Cout));
  float acc = 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    const float* xp = x + ((int64_t)b * Cin + ci) * H * W;
    const float* wp = w + ((int64_t)co * Cin + ci) * K * K;
    for (int kh = 0; kh < K; ++kh) {
      const int hi = ho * stride - pad + kh;
      if (hi < 0 || hi >= H) continue;
      for (int kw = 0; kw < K; ++kw) {
        const int wi = wo * stride - pad + kw;
        if (wi < 0 || wi >= W) continue;
        acc = fmaf(xp[(int64_t)hi * W + wi], wp[kh * K + kw], acc);
      }
    }
  }
  out[idx] = fmaxf(fmaf(acc, scale[co], shift[co]), 0.f);
}

// torch.nn.GRU (1 layer, batch_first) over x[b][t][I] for t < len[b]; out[b][:] = hidden state after the last valid
// step.  Gate order of weight_ih / weight_hh: (r, z, n).  One block per utterance, one thread per hidden unit.
__global__ void gru_last_state_kernel(const float* __restrict__ x, const int64_t* __restrict__ lens, int T, int I, int Hn,
                                      const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                      const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                      float* __restrict__ out) {
  extern __shared__ float sm[];  // h[Hn] | x_t[I]
  float* hs = sm;
  float* xs = sm + Hn;
  const int b = blockIdx.x, j = threadIdx.x;
  int n = lens ? (int)lens[b] : T;
  n = n < 1 ? 1 : (n > T ? T : n);
  if (j < Hn) hs[j] = 0.f;
  __syncthreads();
  for (int t = 0; t < n; ++t) {
    for (int i = threadIdx.x; i < I; i += blockDim.x) xs[i] = x[((int64_t)b * T + t) * I + i];
    __syncthreads();
    float hn = 0.f;
    if (j < Hn) {
      float gi[3], gh[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const float* wi = w_ih + (int64_t)(g * Hn + j) * I;
        const float* wh = w_hh + (int64_t)(g * Hn + j) * Hn;
        float a = b_ih[g * Hn + j], c = b_hh[g * Hn + j];
        for (int i = 0; i < I; ++i) a = fmaf(wi[i], xs[i], a);
        for (int i = 0; i < Hn; ++i) c = fmaf(wh[i], hs[i], c);
        gi[g] = a;
        gh[g] = c;
      }
      const float r = 1.f / (1.f + expf(-(gi[0] + gh[0])));
      const float z = 1.f / (1.f + expf(-(gi[1] + gh[1])));
      const float nn = tanhf(gi[2] + r * gh[2]);
      hn = (1.f - z) * nn + z * hs[j];
    }
    __syncthreads();
    if (j < Hn) hs[j] = hn;
    __syncthreads();
  }
  if (j < Hn) out[(int64_t)b * Hn + j] = hs[j];
}

// StyleTokenLayer + MultiHeadedAttention (style_encoder.py:106-171): q = Wq ref + bq; k/v = Wk/Wv tanh(gst) + b;
// per head softmax(q.k / sqrt(n_feat)) . v; out = Wo o + bo.  One block per utterance, one thread per feature.
__global__ void style_token_attention_kernel(const float* __restrict__ ref, int R, const float* __restrict__ gst, int Tk,
                                             int Dk, int heads, int Fdim, const float* __restrict__ wq,
                                             const float* __restrict__ bq, const float* __restrict__ wk,
                                             const float* __restrict__ bk, const float* __restrict__ wv,
                                             const float* __restrict__ bv, const float* __restrict__ wo,
                                             const float* __restrict__ bo, float* __restrict__ out) {
  extern __shared__ float sm[];  // q[F] | k[Tk][F] | v[Tk][F] | score[heads][Tk] | o[F]
  float* q = sm;
  float* k = q + Fdim;
  float* v = k + Tk * Fdim;
  float* sc = v + Tk * Fdim;
  float* o = sc + heads * Tk;
  const int b = blockIdx.x, f = threadIdx.x;
  const int dh = Fdim / heads;
  if (f < Fdim) {
    float a = bq[f];
    for (int i = 0; i < R; ++i) a = fmaf(wq[(int64_t)f * R + i], ref[(int64_t)b * R + i], a);
    q[f] = a;
    for (int t = 0; t < Tk; ++t) {
      float kk = bk[f], vv = bv[f];
      for (int i = 0; i < Dk; ++i) {
        const float g = tanhf(gst[t * Dk + i]);
        kk = fmaf(wk[(int64_t)f * Dk + i], g, kk);
        vv = fmaf(wv[(int64_t)f * Dk + i], g, vv);
      }
      k[t * Fdim + f] = kk;
      v[t * Fdim + f] = vv;
    }
  }
  __syncthreads();
  if (f < heads * Tk) {
    const int hd = f / Tk, t = f - hd * Tk;
    float s = 0.f;
    for (int i = 0; i < dh; ++i) s = fmaf(q[hd * dh + i], k[t * Fdim + hd * dh + i], s);
    sc[f] = s / sqrtf((float)Fdim);  // math.sqrt(self.d_k * self.h)
  }
  __syncthreads();
  if (f < heads) {
    float mx = -3.4e38f;
    for (int t = 0; t < Tk; ++t) mx = fmaxf(mx, sc[f * Tk + t]);
    float sum = 0.f;
    for (int t = 0; t < Tk; ++t) {
      const float e = expf(sc[f * Tk + t] - mx);
      sc[f * Tk + t] = e;
      sum += e;
    }
    for (int t = 0; t < Tk; ++t) sc[f * Tk + t] /= sum;
  }
  __syncthreads();
  if (f < Fdim) {
    const int hd = f / dh;
    float a = 0.f;
    for (int t = 0; t < Tk; ++t) a = fmaf(sc[hd * Tk + t], v[t * Fdim + f], a);
    o[f] = a;
  }
  __syncthreads();
  if (f < Fdim) {
    float a = bo[f];
    for (int i = 0; i < Fdim; ++i) a = fmaf(wo[(int64_t)f * Fdim + i], o[i], a);
    out[(int64_t)b * Fdim + f] = a;
  }
}

}  // namespace
}  // namespace pttspp

using namespace pttspp;

extern "C" int pttspp_conv2d_bn_relu(const float* x, const float* w, const float* bn_scale, const float* bn_shift,
                                     float* out, int B, int Cin, int H, int W, int Cout, int K, int stride, int pad,
                                     pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(x && w && bn_scale && bn_shift && out, "null argument");
  PT_CHECK(B >= 1 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1 && K >= 1 && stride >= 1 && pad >= 0, "conv2d: bad shape");
  const int Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1;
  PT_CHECK(Ho >= 1 && Wo >= 1, "conv2d: empty output");
  const int64_t total = (int64_t)B * Cout * Ho * Wo;
  ProfScope prof(PROF_OTHER, (cudaStream_t)stream, 2.0 * total * Cin * K * K, 0.0);
  conv2d_bn_relu_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
      x, w, bn_scale, bn_shift, out, B, Cin, H, W, Cout, Ho, Wo, K, stride, pad);
  PT_LAUNCHED();
  PT_API_END
}

extern "C" int pttspp_gru_last_state(const float* x, const int64_t* lens, int B, int T, int I, int Hn, const float* w_ih,
                                     const float* w_hh, const float* b_ih, const float* b_hh, float* out,
                                     pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(x && w_ih && w_hh && b_ih && b_hh && out, "null argument");
  PT_CHECK(B >= 1 && T >= 1 && I >= 1 && Hn >= 1 && Hn <= 1024, "gru: bad shape (hidden <= 1024)");
  const int threads = round_up(Hn, 32);
  const size_t sm = (size_t)(Hn + I) * sizeof(float);
  PT_CHECK(sm <= 48 * 1024, "gru: input size too large");
  gru_last_state_kernel<<<B, threads, sm, (cudaStream_t)stream>>>(x, lens, T, I, Hn, w_ih, w_hh, b_ih, b_hh, out);
  PT_LAUNCHED();
  PT_API_END
}

extern "C" int pttspp_style_token_attention(const float* ref, int B, int R, const float* gst_embs, int Tk, int Dk, int heads,
                                            int Fdim, const float* wq, const float* bq, const float* wk, const float* bk,
                                            const float* wv, const float* bv, const float* wo, const float* bo, float* out,
                                            pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(ref && gst_embs && wq && bq && wk && bk && wv && bv && wo && bo && out, "null argument");
  PT_CHECK(B >= 1 && R >= 1 && Tk >= 1 && Dk >= 1 && heads >= 1 && Fdim % heads == 0 && Fdim <= 1024 && heads * Tk <= Fdim,
           "style token attention: bad shape");
  const size_t sm = (size_t)(2 * Fdim + 2 * Tk * Fdim + heads * Tk) * sizeof(float);
  PT_CHECK(sm <= 48 * 1024, "style token attention: too many tokens");
  style_token_attention_kernel<<<B, round_up(Fdim, 32), sm, (cudaStream_t)stream>>>(ref, R, gst_embs, Tk, Dk, heads, Fdim, wq,
                                                                                   bq, wk, bk, wv, bv, wo, bo, out);
  PT_LAUNCHED();
  PT_API_END
}

// ---- BERT building blocks (the prompt encoder's sentence embedding) ----------------------
// BertEmbeddings (HF transformers modeling_bert.py, the reference's dependency at promptttspp/modules/prompt_encoder.py:
// 25,33-38): word + position + token-type(0) embeddings (the LayerNorm that follows is pttspp_layernorm_cl);
// BertSelfAttention: softmax(q k^T / sqrt(dk) + (1 - mask) * finfo.min) v on a fused [B][T][3*H*dk] q|k|v buffer.
namespace pttspp {
namespace {

__global__ void __launch_bounds__(256) bert_embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                                         const float* __restrict__ pos, const float* __restrict__ type0,
                                                         float* __restrict__ out, int T, int Hd, int vocab, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Hd);
  const int64_t bt = i / Hd;
  const int t = (int)(bt % T);
  int64_t id = ids[bt];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  out[i] = (word[id * Hd + c] + type0[c]) + pos[(int64_t)t * Hd + c];  // inputs_embeds + token_type, then + position
}

// one warp per query row; scores of that row in shared memory
__global__ void __launch_bounds__(128) mha_masked_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ mask,
                                                         float* __restrict__ out, int T, int heads, int dk) {
  extern __shared__ float sc[];  // [4 warps][T]
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hd = heads * dk, ld = 3 * Hd;
  const float* base = qkv + (int64_t)b * T * ld;
  float* s = sc + warp * T;
  const float scale = 1.f / sqrtf((float)dk);
  for (int qi = blockIdx.x * 4 + warp; qi < T; qi += gridDim.x * 4) {
    const float* q = base + (int64_t)qi * ld + h * dk;
    for (int j = 0; j < T; ++j) {
      const float* k = base + (int64_t)j * ld + Hd + h * dk;
      float a = 0.f;
      for (int d = lane; d < dk; d += 32) a = fmaf(q[d], k[d], a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) s[j] = a * scale + ((mask && mask[(int64_t)b * T + j] == 0) ? -3.4028234663852886e38f : 0.f);
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, s[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = expf(s[j] - mx);
      s[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    for (int d = lane; d < dk; d += 32) {
      float a = 0.f;
      for (int j = 0; j < T; ++j) a = fmaf(s[j], base[(int64_t)j * ld + 2 * Hd + h * dk + d], a);
      out[((int64_t)b * T + qi) * Hd + h * dk + d] = a / sum;
    }
    __syncwarp();
  }
}

}  // namespace
}  // namespace pttspp

extern "C" int pttspp_bert_embed(const int64_t* ids, int B, int T, const float* word_emb, int vocab, const float* pos_emb,
                                 const float* type_emb0, int hidden, float* out, pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(ids && word_emb && pos_emb && type_emb0 && out, "null argument");
  PT_CHECK(B >= 1 && T >= 1 && hidden >= 1 && vocab >= 1, "bert_embed: bad shape");
  const int64_t total = (int64_t)B * T * hidden;
  bert_embed_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(ids, word_emb, pos_emb, type_emb0,
                                                                                       out, T, hidden, vocab, total);
  PT_LAUNCHED();
  PT_API_END
}

extern "C" int pttspp_mha_masked(const float* qkv, const int64_t* key_mask, int B, int T, int heads, int dk, float* out,
                                 pttspp_stream_t stream) {
  PT_API_BEGIN
  PT_CHECK(qkv && out, "null argument");
  PT_CHECK(B >= 1 && T >= 1 && heads >= 1 && dk >= 1 && T <= 2048, "mha: bad shape (T <= 2048)");
  ProfScope prof(PROF_ATTENTION, (cudaStream_t)stream, 4.0 * B * heads * (double)T * T * dk, 0.0);
  dim3 grid(std::min(ceil_div(T, 4), 64), heads, B);
  mha_masked_kernel<<<grid, 128, (size_t)4 * T * sizeof(float), (cudaStream_t)stream>>>(qkv, key_mask, out, T, heads, dk);
  PT_LAUNCHED();
  PT_API_END
}
