// common.cuh — error plumbing and small device helpers shared by the libvqvae_b200 translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "vqb.h"

namespace vqb {

// thread-local last-error text (vqb_last_error)
char* err_buf();
int set_err(int code, const char* fmt, ...);
// VQB_OK if the current device is sm_100-class (cached per device)
int require_arch();
// bumps the process-wide kernel-launch counter (vqb_kernel_launch_count)
void count_launch();
// cancels the count_launch() of a caller whose kernel was queued for a batched launch instead of launched (reduce_begin)
void uncount_launch();

#define VQB_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::vqb::set_err(VQB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define VQB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::vqb::set_err(VQB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                            \
  } while (0)

#define VQB_LAUNCH_CHECK()                                                                       \
  do {                                                                                           \
    ::vqb::count_launch();                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess)                                                                      \
      return ::vqb::set_err(VQB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                 \
  } while (0)

#define VQB_ARCH()                        \
  do {                                    \
    int a__ = ::vqb::require_arch();      \
    if (a__ != VQB_OK) return a__;        \
  } while (0)

// fixed-order reduction of per-chunk partial sums (conv_fp32.cu)
void reduce_chunks_strided(const float* partial, int nchunk, long stride, int offset, int n, float* out, cudaStream_t st);
void reduce_begin();
int reduce_flush(cudaStream_t st);
// tensor-core weight gradient (wgrad_tc.cu)
bool wgrad_tc_supported(const vqb_conv_desc* d);
size_t wgrad_tc_workspace_bytes(const vqb_conv_desc* d);
int conv1d_wgrad_tc(const vqb_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t st);
size_t resblock_wgrad_tc_workspace_bytes(const vqb_conv_desc* d, int n);
int resblock_wgrad_tc(const vqb_conv_desc* d, int n, const int* dilations, const float* const* x, const float* const* h,
                      const float* const* dy, const float* const* dh, float* const* dw1, float* const* db1,
                      float* const* dw2, float* const* db2, void* ws, size_t ws_bytes, cudaStream_t st);

// tensor-core stride-2 convolutions (conv_tc.cu)
// conv3_tc.cu: k = 3, stride-1 convolutions 32 <-> 64 channels at latent rate (forward and data gradient)
bool conv3_tc_supported(const vqb_conv_desc* d);
int conv3_fwd_tc(const vqb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int conv3_dgrad_tc(const vqb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st);
bool conv_tc_supported(const vqb_conv_desc* d);
int conv1d_fwd_tc(const vqb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int conv1d_dgrad_tc(const vqb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st);
int conv1d_transpose_fwd_tc(const vqb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int conv1d_transpose_dgrad_tc(const vqb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st);

// tensor-core weight gradient of the k = 4, stride-2, 32 <-> 32 convolutions (wgrad4_tc.cu)
size_t wgrad4_tc_workspace_bytes(int B, int Lo);
int wgrad4_tc(int precision, const float* ga, int Lg, const float* ot, int Lo, int B, float* dw, float* dbias, bool bias_from_ga,
              void* ws, size_t ws_bytes, cudaStream_t st);

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// The ~600 kernels of a training step form one dependency chain, and the persistent tensor-core kernels have a prologue that
// does not depend on the previous kernel (TMEM allocation, mbarrier init, packing the layer's weights into shared-memory
// operand images).  Launched with the programmatic-stream-serialization attribute, a kernel may become resident while its
// predecessor is still draining; it calls pdl_wait() before its first access to anything the predecessor wrote (and before its
// first global write), and pdl_launch_dependents() once its own prologue is done so that ITS successor can do the same.
// Rules kept by every kernel launched through launch_pdl: (1) every CTA executes pdl_wait() before exiting (completion is
// transitive only then); (2) nothing read before pdl_wait() is written by a kernel of the same graph launch (weights are
// written by Adam in the second graph of a step; layer inputs never are).  VQB_PDL=0 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// tier 1: the persistent tensor-core kernels (prologue worth overlapping); tier 2: plain kernels that wait first thing.
// VQB_PDL = highest tier that gets the attribute (0 = none).
bool pdl_enabled(int tier);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_tier(int tier, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled(tier) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#define launch_pdl(...) launch_pdl_tier(1, __VA_ARGS__)
#define launch_pdl2(...) launch_pdl_tier(2, __VA_ARGS__)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block sum (fixed tree); result valid in thread 0.  `red` has >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = 0.f;
  if (wid == 0) {
    r = lane < nw ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

}  // namespace vqb
