// Library-wide state (last error, launch counter), the host tensor store behind the
// *_set_tensor entry points, and small helpers shared by the model-level handles.
#include "common.h"

namespace pttspp {

static thread_local std::string g_last_error;
thread_local int64_t g_launch_count = 0;

void set_last_error(const std::string& msg) { g_last_error = msg; }

// ---- per-launch profiler ---------------------------------------------------------------------
namespace {
struct ProfRec {
  int tag;
  cudaEvent_t e0, e1;
};
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
double g_prof_flops[PROF_NUM_TAGS] = {0}, g_prof_bytes[PROF_NUM_TAGS] = {0};
int64_t g_prof_calls[PROF_NUM_TAGS] = {0};
}  // namespace

ProfScope::ProfScope(int tag_, cudaStream_t s_, double flops, double bytes) : tag(tag_), s(s_) {
  if (!g_prof_on) return;
  g_prof_flops[tag] += flops;
  g_prof_bytes[tag] += bytes;
  g_prof_calls[tag] += 1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
}
ProfScope::~ProfScope() {
  if (!e0) return;
  cudaEventRecord(e1, s);
  g_prof_recs.push_back({tag, e0, e1});
}

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
void* DeviceBuffers::upload_bytes(const void* host, size_t bytes) {
  void* p = nullptr;
  PT_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
  ptrs.push_back(p);
  if (bytes) PT_CUDA(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
  return p;
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

void attach_split_weights(const TensorStore& st, DeviceBuffers& dev, const std::string& prefix, PackedConv& c,
                          bool interleave_halves) {
  const int64_t n = (int64_t)c.Cout * c.Cin * c.K;
  std::vector<uint16_t> hi((size_t)n), lo((size_t)n);
  if (st.has(prefix + ".weight")) {
    pack_conv_weight_split(st.get(prefix + ".weight", n).data.data(), nullptr, c.Cout, c.Cin, c.K, hi.data(), lo.data(),
                           interleave_halves, &c.w_scale_inv);
  } else {
    pack_conv_weight_split(st.get(prefix + ".weight_v", n).data.data(), st.get(prefix + ".weight_g", c.Cout).data.data(),
                           c.Cout, c.Cin, c.K, hi.data(), lo.data(), interleave_halves, &c.w_scale_inv);
  }
  c.w_hi = dev.upload_bytes(hi.data(), hi.size() * 2);
  c.w_lo = dev.upload_bytes(lo.data(), lo.size() * 2);
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

extern "C" void pttspp_prof_enable(int on) {
  using namespace pttspp;
  for (auto& r : g_prof_recs) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof_recs.clear();
  for (int i = 0; i < PROF_NUM_TAGS; ++i) g_prof_flops[i] = g_prof_bytes[i] = 0.0, g_prof_calls[i] = 0;
  g_prof_on = on != 0;
}

extern "C" int pttspp_prof_report(double* ms, double* flops, double* bytes, int64_t* calls, int ntags) {
  PT_API_BEGIN
  using namespace pttspp;
  PT_CHECK(ms && flops && bytes && calls && ntags >= PROF_NUM_TAGS, "prof_report: need %d slots", PROF_NUM_TAGS);
  PT_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < PROF_NUM_TAGS; ++i) ms[i] = 0.0, flops[i] = g_prof_flops[i], bytes[i] = g_prof_bytes[i], calls[i] = g_prof_calls[i];
  for (auto& r : g_prof_recs) {
    float t = 0.f;
    PT_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms[r.tag] += t;
  }
  PT_API_END
}

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
