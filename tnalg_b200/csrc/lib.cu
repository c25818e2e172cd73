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

// register-resident DMMA.8x8x4 stream (16 independent accumulators per warp, 8 warps per CTA, 2 CTAs per SM): the issue-rate
// ceiling of the FP64 tensor pipe, measured on the device the caller runs on (bench.py quotes it next to the cuBLAS DGEMM rate)
namespace tn {
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5 + threadIdx.x * 1e-9;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace tn

extern "C" int tn_measure_dmma_peak(double* tflops_out, void* scratch, size_t scratch_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = tn::sm_count() * 2, iters = 8192;
  TN_REQUIRE(tflops_out && scratch && scratch_bytes >= sizeof(double) * (size_t)grid * 256, "tn_measure_dmma_peak: scratch of %zu bytes needed",
             sizeof(double) * (size_t)grid * 256);
  cudaEvent_t e0, e1;
  TN_CUDA(cudaEventCreate(&e0));
  TN_CUDA(cudaEventCreate(&e1));
  tn::dmma_peak_kernel<<<grid, 256, 0, stream>>>(static_cast<double*>(scratch), 64, 1.0);
  TN_LAUNCHED();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    TN_CUDA(cudaEventRecord(e0, stream));
    tn::dmma_peak_kernel<<<grid, 256, 0, stream>>>(static_cast<double*>(scratch), iters, 1.0);
    TN_LAUNCHED();
    TN_CUDA(cudaEventRecord(e1, stream));
    TN_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    TN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = 2.0 * 256 * 16 * (double)iters * ((double)grid * 8) / best / 1e9;
  return TN_OK;
}

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
