// Log-mel spectrogram of a waveform batch: promptttspp/transforms/mel.py:18-34 (torchaudio MelSpectrogram with the kwargs
// of conf/transforms/mel.yaml: n_fft 512, win 480, hop 240, power 1, slaney mel scale + norm, center / reflect), used by
// app.py:93-100 and egs/proposed/bin/synthesize.py:172-175 to turn a reference utterance into the style encoder's input.
//
// One CTA per (frame, utterance): the windowed frame (reflect-padded at the signal's ends, like torch.stft(center=True))
// sits in shared memory next to one period of the twiddle factors; thread k evaluates DFT bin k directly (n_fft real
// multiply-adds per component -- 0.26 MFLOP per frame, the whole of cfg3's audio is < 10 GFLOP, so the transform is bound
// by its single pass over the waveform and not worth an FFT butterfly network), the magnitudes stay in shared memory
// and the first n_mels threads contract them with the (sparse, triangular) filterbank.  The waveform is read once, only
// the n_mels log-energies per frame are written (and the linear spectrogram when the caller asks for it).
#include "common.h"

namespace pttspp {
namespace {

constexpr int MEL_MAX_NFFT = 2048;

__global__ void __launch_bounds__(256) mel_spectrogram_kernel(const float* __restrict__ wav, int L, int n_fft, int hop,
                                                              int frames, const float* __restrict__ window,
                                                              const float* __restrict__ fb, int n_mels, int power2,
                                                              float log_floor, float* __restrict__ spec_out,
                                                              float* __restrict__ mel_out) {
  extern __shared__ float sm[];
  float* xs = sm;                 // [n_fft] windowed frame
  float* cs = xs + n_fft;         // [n_fft] cos(2 pi j / n_fft)
  float* sn = cs + n_fft;         // [n_fft] sin(2 pi j / n_fft)
  float* mag = sn + n_fft;        // [n_fft / 2 + 1]
  const int t = blockIdx.x, b = blockIdx.y;
  const int n_freq = n_fft / 2 + 1;
  const float* w = wav + (int64_t)b * L;
  const int start = t * hop - n_fft / 2;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    int j = start + n;
    // reflect padding without repeating the edge sample (torch.nn.functional.pad(mode="reflect")); L > n_fft / 2
    if (j < 0) j = -j;
    if (j >= L) j = 2 * (L - 1) - j;
    xs[n] = w[j] * window[n];
    float s, c;
    sincospif(2.0f * (float)n / (float)n_fft, &s, &c);
    cs[n] = c;
    sn[n] = s;
  }
  __syncthreads();
  const int mask = n_fft - 1;
  for (int k = threadIdx.x; k < n_freq; k += blockDim.x) {
    float re0 = 0.f, im0 = 0.f, re1 = 0.f, im1 = 0.f;
    int idx = 0;
    for (int n = 0; n < n_fft; n += 2) {
      re0 = fmaf(xs[n], cs[idx], re0);
      im0 = fmaf(xs[n], sn[idx], im0);
      idx = (idx + k) & mask;
      re1 = fmaf(xs[n + 1], cs[idx], re1);
      im1 = fmaf(xs[n + 1], sn[idx], im1);
      idx = (idx + k) & mask;
    }
    const float re = re0 + re1, im = im0 + im1;
    const float p = re * re + im * im;
    const float v = power2 ? p : sqrtf(p);
    mag[k] = v;
    if (spec_out) spec_out[((int64_t)b * n_freq + k) * frames + t] = v;
  }
  __syncthreads();
  if (mel_out) {
    for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < n_freq; ++k) acc = fmaf(mag[k], __ldg(fb + (int64_t)k * n_mels + m), acc);
      mel_out[((int64_t)b * n_mels + m) * frames + t] = logf(fmaxf(acc, log_floor));
    }
  }
}

__global__ void __launch_bounds__(128) mel_from_spec_kernel(const float* __restrict__ spec, int n_freq, int frames,
                                                            const float* __restrict__ fb, int n_mels, float log_floor,
                                                            float* __restrict__ mel_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y, b = blockIdx.z;
  if (t >= frames) return;
  const float* sp = spec + (int64_t)b * n_freq * frames + t;
  float acc = 0.f;
  for (int k = 0; k < n_freq; ++k) acc = fmaf(sp[(int64_t)k * frames], __ldg(fb + (int64_t)k * n_mels + m), acc);
  mel_out[((int64_t)b * n_mels + m) * frames + t] = logf(fmaxf(acc, log_floor));
}

}  // namespace
}  // namespace pttspp

extern "C" int pttspp_mel_spectrogram(const float* wav, int B, int L, int n_fft, int hop, const float* window,
                                      const float* fb, int n_mels, int power, float log_floor, float* spec_out,
                                      float* mel_out, pttspp_stream_t stream) {
  PT_API_BEGIN
  using namespace pttspp;
  PT_CHECK(wav && window && (spec_out || mel_out), "mel_spectrogram: null argument");
  PT_CHECK(!mel_out || (fb && n_mels >= 1), "mel_spectrogram: the mel output needs a filterbank");
  PT_CHECK(n_fft >= 32 && n_fft <= MEL_MAX_NFFT && (n_fft & (n_fft - 1)) == 0, "mel_spectrogram: n_fft must be a power "
           "of two in [32, %d] (got %d)", MEL_MAX_NFFT, n_fft);
  PT_CHECK(hop >= 1 && (power == 1 || power == 2), "mel_spectrogram: hop >= 1 and power in {1, 2} are supported");
  PT_CHECK(B >= 1 && L > n_fft / 2, "mel_spectrogram: reflect padding needs more than n_fft/2 = %d samples (got %d)",
           n_fft / 2, L);
  const int frames = 1 + L / hop;
  const size_t smem = (size_t)(3 * n_fft + n_fft / 2 + 1) * sizeof(float);
  dim3 grid(frames, B);
  mel_spectrogram_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(wav, L, n_fft, hop, frames, window, fb, n_mels,
                                                                    power == 2, log_floor, spec_out, mel_out);
  PT_LAUNCHED();
  PT_API_END
}

extern "C" int pttspp_mel_from_spec(const float* spec, int B, int n_freq, int frames, const float* fb, int n_mels,
                                    float log_floor, float* mel_out, pttspp_stream_t stream) {
  PT_API_BEGIN
  using namespace pttspp;
  PT_CHECK(spec && fb && mel_out && B >= 1 && n_freq >= 1 && frames >= 1 && n_mels >= 1, "mel_from_spec: bad argument");
  dim3 grid(ceil_div(frames, 128), n_mels, B);
  mel_from_spec_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(spec, n_freq, frames, fb, n_mels, log_floor, mel_out);
  PT_LAUNCHED();
  PT_API_END
}
