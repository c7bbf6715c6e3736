// Library-wide state (last error, launch counter), the host tensor store behind the
// *_set_tensor entry points, and small helpers shared by the model-level handles.
#include "common.h"

namespace pttspp {

static thread_local std::string g_last_error;
thread_local int64_t g_launch_count = 0;

void set_last_error(const std::string& msg) { g_last_error = msg; }

void TensorStore::set(const char* name, const float* data, const int64_t* shape, int ndim, cudaStream_t s) {
  PT_CHECK(name && data && (shape || ndim == 0), "set_tensor: null argument");
  HostTensor ht;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    PT_CHECK(shape[i] >= 0, "set_tensor(%s): negative dim", name);
    ht.shape.push_back(shape[i]);
    n *= shape[i];
  }
  ht.data.resize((size_t)n);
  if (n > 0) {
    // `data` may live on the host or on the device (unified virtual addressing sorts it out)
    PT_CUDA(cudaMemcpyAsync(ht.data.data(), data, (size_t)n * sizeof(float), cudaMemcpyDefault, s));
    PT_CUDA(cudaStreamSynchronize(s));
  }
  t[name] = std::move(ht);
}

const HostTensor& TensorStore::get(const std::string& name) const {
  auto it = t.find(name);
  PT_CHECK(it != t.end(), "missing tensor \"%s\" (not provided through set_tensor)", name.c_str());
  return it->second;
}

const HostTensor& TensorStore::get(const std::string& name, int64_t numel) const {
  const HostTensor& h = get(name);
  PT_CHECK(h.numel() == numel, "tensor \"%s\" has %lld elements, expected %lld", name.c_str(),
           (long long)h.numel(), (long long)numel);
  return h;
}

DeviceBuffers::~DeviceBuffers() { release(); }
void DeviceBuffers::release() {
  for (void* p : ptrs) cudaFree(p);
  ptrs.clear();
}
float* DeviceBuffers::upload(const float* host, size_t n) {
  void* p = nullptr;
  PT_CUDA(cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(float)));
  ptrs.push_back(p);
  if (n) PT_CUDA(cudaMemcpy(p, host, n * sizeof(float), cudaMemcpyHostToDevice));
  return (float*)p;
}
float* DeviceBuffers::upload(const std::vector<float>& host) { return upload(host.data(), host.size()); }

PackedConv load_conv1d(const TensorStore& st, DeviceBuffers& dev, const std::string& prefix, int Cout, int Cin, int K,
                       int dil, int pad, bool interleave_halves, bool has_bias) {
  PackedConv c;
  c.Cin = Cin; c.Cout = Cout; c.K = K; c.dil = dil; c.pad = pad;
  c.w_ld = round_up(Cout, 4);
  std::vector<float> packed((size_t)K * Cin * c.w_ld);
  const int64_t n = (int64_t)Cout * Cin * K;
  if (st.has(prefix + ".weight")) {
    pack_conv_weight(st.get(prefix + ".weight", n).data.data(), nullptr, Cout, Cin, K, packed.data(), c.w_ld,
                     interleave_halves, 0);
  } else {
    const HostTensor& v = st.get(prefix + ".weight_v", n);
    const HostTensor& g = st.get(prefix + ".weight_g", Cout);
    pack_conv_weight(v.data.data(), g.data.data(), Cout, Cin, K, packed.data(), c.w_ld, interleave_halves, 0);
  }
  c.w = dev.upload(packed);
  if (has_bias) {
    const HostTensor& b = st.get(prefix + ".bias", Cout);
    std::vector<float> bb(b.data);
    if (interleave_halves) {
      const int half = Cout / 2;
      for (int co = 0; co < Cout; ++co) bb[co < half ? 2 * co : 2 * (co - half) + 1] = b.data[co];
    }
    c.bias = dev.upload(bb);
  }
  return c;
}

PackedConv load_linear(const TensorStore& st, DeviceBuffers& dev, const std::string& prefix, int Cout, int Cin,
                       bool has_bias) {
  return load_conv1d(st, dev, prefix, Cout, Cin, 1, 1, 0, false, has_bias);
}

pttspp_conv1d_desc conv_desc(const PackedConv& c, const float* in, int B, int T, float* out) {
  pttspp_conv1d_desc d;
  memset(&d, 0, sizeof(d));
  d.in = in; d.in_bs = (int64_t)T * c.Cin; d.in_ld = c.Cin; d.T_in = T; d.Cin = c.Cin;
  d.w = c.w; d.w_ld = c.w_ld; d.bias = c.bias;
  d.K = c.K; d.dil = c.dil; d.pad = c.pad; d.in_stride = 1;
  d.out = out; d.out_bs = (int64_t)T * c.Cout; d.out_ld = c.Cout; d.T_out = T; d.Cout = c.Cout;
  d.m_begin = 0; d.M = T; d.out_mul = 1; d.out_off = 0;
  d.act = PTTSPP_ACT_NONE; d.acc_scale = 1.f; d.res_scale = 1.f; d.alpha = 1.f; d.beta = 0.f;
  d.B = B;
  return d;
}

}  // namespace pttspp

extern "C" const char* pttspp_last_error(void) { return pttspp::g_last_error.c_str(); }
extern "C" int pttspp_abi_version(void) { return PTTSPP_ABI_VERSION; }
extern "C" int64_t pttspp_launch_count(void) { return pttspp::g_launch_count; }
extern "C" void pttspp_reset_launch_count(void) { pttspp::g_launch_count = 0; }
extern "C" int pttspp_device_check(void) {
  PT_API_BEGIN
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PT_CUDA(cudaGetDeviceProperties(&prop, dev));
  PT_CHECK(prop.major == 10, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev, prop.name,
           prop.major, prop.minor);
  PT_API_END
}
