// Library-level plumbing: error text, version, device query, launch counter.
#include "common.cuh"

namespace tn {
static thread_local std::string g_err;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

static std::atomic<int> g_deterministic{0};
bool deterministic_mode() { return g_deterministic.load() != 0; }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}
}  // namespace tn

extern "C" const char* tn_last_error(void) { return tn::g_err.c_str(); }
extern "C" int tn_version(void) { return 200; }
extern "C" int tn_set_deterministic(int on) {
  int old = tn::g_deterministic.exchange(on ? 1 : 0);
  return old;
}
extern "C" long long tn_launch_count(void) { return tn::g_launches.load(); }
extern "C" void tn_launch_count_reset(void) { tn::g_launches.store(0); }

extern "C" int tn_device_info(int* sms, int* major, int* minor) {
  int dev = 0, a = 0, b = 0, c = 0;
  TN_CUDA(cudaGetDevice(&dev));
  TN_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  TN_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  TN_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (major) *major = b;
  if (minor) *minor = c;
  if (b != 10) {
    tn::set_error("tnalg_b200 is built for sm_100a only; device has compute capability %d.%d", b, c);
    return TN_ERR_DEVICE;
  }
  return TN_OK;
}
