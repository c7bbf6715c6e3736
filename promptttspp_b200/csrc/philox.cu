// Standard-normal draws that reproduce torch's CUDA generator stream bit for bit, so the sampler can draw the per-step
// noise of diffusion.py:218 (`noise_like(x.shape)` -> torch.randn) INSIDE the native loop, one [B][mel][Ty] buffer at a
// time, instead of the caller materialising all K_step draws up front (1.3 GB at cfg2).
// Recipe (ATen/native/cuda/DistributionTemplates.h, `normal_` on a float CUDA tensor): block 256, grid = min(#SM *
// (max threads per SM / 256), ceil(numel / 256)), every thread runs curand_init(seed, thread index, offset) on a
// Philox4_32_10 state and per grid-stride iteration takes one curand_normal4, writing component ii to element
// linear_index + ii * (block * grid); the generator's offset then advances by
// ((numel - 1) / (block * grid * 4) + 1) * 4.  cuRAND's device headers provide the generator itself.
#include <curand_kernel.h>

#include "common.h"

namespace pttspp {
namespace {

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, long long numel,
                                                            unsigned long long seed, unsigned long long offset) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  curandStatePhilox4_32_10_t state;
  curand_init(seed, (unsigned long long)idx, offset, &state);
  const long long stride = (long long)blockDim.x * gridDim.x;
  const long long rounded = ((numel - 1) / (stride * 4) + 1) * stride * 4;
  for (long long li0 = idx; li0 < rounded; li0 += stride * 4) {
    const float4 r = curand_normal4(&state);
    const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long li = li0 + stride * ii;
      if (li < numel) out[li] = v[ii] * 1.0f + 0.0f;  // transformation::normal(rand, mean 0, std 1)
    }
  }
}

}  // namespace

// returns the amount the generator's offset advances (what torch adds to philox_offset_per_thread for this call)
uint64_t philox_normal(float* out, int64_t numel, uint64_t seed, uint64_t offset, cudaStream_t s) {
  PT_CHECK(out && numel >= 1, "philox_normal: bad argument");
  int dev = 0, sms = 0, tpsm = 0;
  PT_CUDA(cudaGetDevice(&dev));
  PT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PT_CUDA(cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
  const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sms * (uint64_t)(tpsm / 256), ((uint64_t)numel + 255) / 256);
  philox_normal_kernel<<<grid, 256, 0, s>>>(out, (long long)numel, seed, offset);
  PT_LAUNCHED();
  return (((uint64_t)numel - 1) / (256ull * grid * 4ull) + 1) * 4ull;
}

}  // namespace pttspp

extern "C" int pttspp_philox_normal(float* out, int64_t numel, uint64_t seed, uint64_t offset, uint64_t* offset_advance,
                                    pttspp_stream_t stream) {
  PT_API_BEGIN
  const uint64_t adv = pttspp::philox_normal(out, numel, seed, offset, (cudaStream_t)stream);
  if (offset_advance) *offset_advance = adv;
  PT_API_END
}
