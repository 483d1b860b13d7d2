// lib.cu — version, error text and the architecture gate of libvqvae_b200.so.
#include <stdlib.h>

#include "common.cuh"
#include <atomic>

#ifndef VQB_PDL_DEFAULT
#define VQB_PDL_DEFAULT 1
#endif

namespace vqb {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void uncount_launch() { g_launches.fetch_sub(1, std::memory_order_relaxed); }

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int require_arch() {
  static thread_local int cached_dev = -1, cached_rc = 0;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_err(VQB_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return set_err(VQB_ERR_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
  cached_dev = dev;
  cached_rc = major == 10 ? VQB_OK
                          : set_err(VQB_ERR_ARCH, "device %d has compute capability %d.x; libvqvae_b200 is sm_100a only", dev, major);
  return cached_rc;
}

bool pdl_enabled(int tier) {  // read per call: tests and A/B measurements flip it with the environment
  const char* e = getenv("VQB_PDL");
  const int level = e ? atoi(e) : VQB_PDL_DEFAULT;
  return tier <= level;
}

}  // namespace vqb

extern "C" {

int vqb_version(void) { return VQB_VERSION; }

int64_t vqb_kernel_launch_count(void) { return (int64_t)vqb::g_launches.load(); }

const char* vqb_last_error(void) { return vqb::err_buf(); }

int vqb_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n)
    return vqb::set_err(VQB_ERR_ARCH, "no CUDA device %d (%s)", device, e != cudaSuccess ? cudaGetErrorString(e) : "out of range");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return vqb::set_err(VQB_ERR_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
  if (major != 10)
    return vqb::set_err(VQB_ERR_ARCH, "device %d has compute capability %d.x; libvqvae_b200 is sm_100a only", device, major);
  return VQB_OK;
}

}  // extern "C"
