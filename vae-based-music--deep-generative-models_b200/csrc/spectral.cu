// spectral.cu — the element-wise halves of the multi-scale spectral convergence loss (vqvae.py:309-326, data_utils.py:25-40)
// around the FFT itself (cuFFT, called by the host layer):
//   vqb_stft_frames      x [B, T] -> frames [B, F, n_fft]: frame f = x[f*hop .. f*hop+win) * periodic Hann, zero padded at
//                        the END up to n_fft (tf.signal.stft, pad_end=False: F = 1 + (T - win) / hop)
//   vqb_spec_mag         spectrum S [B, F, bins] complex -> |S| and, per example, sum |S|^2           (target side, cached)
//   vqb_spec_diff        S_r, |S_t| -> per-example sum (|S_t| - |S_r|)^2 (the squared Frobenius norm of data_utils.norm)
//   vqb_spec_loss        the three scales' sums -> loss = mean_b mean_s sqrt(sum_diff) / sqrt(sum_target)
//   vqb_spec_grad        dL/dS_r as the one-sided spectrum whose irfft (times n_fft) is the gradient of the frames:
//                        g_b * (|S_r| - |S_t|) * S_r / |S_r|, interior bins halved (irfft doubles them), 0 where |S_r| = 0
//   vqb_stft_frames_bwd  overlap-add of the windowed frame gradients back onto [B, T] (gather form: no atomics, fixed order)
// All bandwidth-bound, one pass each; every reduction is a fixed-order tree (deterministic).
#include "common.cuh"

namespace vqb {

// periodic Hann: w[n] = 0.5 - 0.5 cos(2 pi n / win)  (tf.signal.hann_window(periodic=True))
__device__ __forceinline__ float hann(int n, int win) { return 0.5f - 0.5f * cospif(2.0f * (float)n / (float)win); }

__global__ void __launch_bounds__(256) stft_frames_kernel(const float* __restrict__ x, float* __restrict__ frames, int T, int F,
                                                          int n_fft, int hop, int win, long total) {
  const long e = (long)blockIdx.x * 256 + threadIdx.x;  // one thread per 4 consecutive samples of a frame
  const int q4 = n_fft >> 2;
  if (e >= total) return;
  const int n = (int)(e % q4) * 4;
  const long bf = e / q4;
  const int f = (int)(bf % F);
  const long b = bf / F;
  const float* xr = x + b * T + (long)f * hop;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < win) v.x = xr[n] * hann(n, win);
  if (n + 1 < win) v.y = xr[n + 1] * hann(n + 1, win);
  if (n + 2 < win) v.z = xr[n + 2] * hann(n + 2, win);
  if (n + 3 < win) v.w = xr[n + 3] * hann(n + 3, win);
  *reinterpret_cast<float4*>(frames + e * 4) = v;
}

constexpr int SP_CHUNK = 4096;  // spectrum elements per CTA (256 threads x 16)

// per-example partial sums: grid (chunks, B); partial[b * chunks + c]
__global__ void __launch_bounds__(256) spec_mag_kernel(const float2* __restrict__ S, float* __restrict__ mag, long per_ex,
                                                       float* __restrict__ partial) {
  __shared__ float red[32];
  const long b = blockIdx.y;
  const long e0 = (long)blockIdx.x * SP_CHUNK;
  float acc = 0.f;
  for (int k = threadIdx.x; k < SP_CHUNK; k += 256) {
    const long e = e0 + k;
    if (e < per_ex) {
      const float2 s = S[b * per_ex + e];
      const float m = sqrtf(fmaf(s.x, s.x, s.y * s.y));
      mag[b * per_ex + e] = m;
      acc = fmaf(m, m, acc);
    }
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[b * gridDim.x + blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) spec_diff_kernel(const float2* __restrict__ S, const float* __restrict__ mag_t, long per_ex,
                                                        float* __restrict__ partial) {
  __shared__ float red[32];
  const long b = blockIdx.y;
  const long e0 = (long)blockIdx.x * SP_CHUNK;
  float acc = 0.f;
  for (int k = threadIdx.x; k < SP_CHUNK; k += 256) {
    const long e = e0 + k;
    if (e < per_ex) {
      const float2 s = S[b * per_ex + e];
      const float d = mag_t[b * per_ex + e] - sqrtf(fmaf(s.x, s.x, s.y * s.y));
      acc = fmaf(d, d, acc);
    }
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) partial[b * gridDim.x + blockIdx.x] = t;
}

// sums[b] = sum_c partial[b * chunks + c] (fixed order); one warp per example
__global__ void __launch_bounds__(256) spec_sum_kernel(const float* __restrict__ partial, int chunks, int B, float* __restrict__ sums) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float s = 0.f;
  for (int c = lane; c < chunks; c += 32) s += partial[(long)b * chunks + c];
  s = warp_sum(s);
  if (lane == 0) sums[b] = s;
}

// loss[0] = mean_b mean_s sqrt(d_s[b]) / sqrt(t_s[b]); coef[s * B + b] = 1 / (nscales * B * sqrt(t_s[b]) * sqrt(d_s[b])) (0 if d = 0)
__global__ void __launch_bounds__(256) spec_loss_kernel(const float* __restrict__ dsum, const float* __restrict__ tsum, int nscales, int B,
                                                        float* __restrict__ loss, float* __restrict__ coef) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int e = threadIdx.x; e < nscales * B; e += 256) {
    const float nd = sqrtf(dsum[e]), nt = sqrtf(tsum[e]);
    acc += nd / nt;
    if (coef) coef[e] = nd > 0.f ? 1.0f / ((float)(nscales * B) * nt * nd) : 0.f;
  }
  const float t = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = t / (float)(nscales * B);
}

// G[b,f,k] = upstream * coef[b] * (|S| - |S_t|) * S / |S| * (interior ? 0.5 : 1) * n_fft   (irfft divides by n_fft and doubles the interior)
__global__ void __launch_bounds__(256) spec_grad_kernel(const float2* __restrict__ S, const float* __restrict__ mag_t,
                                                        const float* __restrict__ coef, const float* __restrict__ upstream, long per_ex,
                                                        int bins, float scale_n, float2* __restrict__ G, long total) {
  const long e = (long)blockIdx.x * 256 + threadIdx.x;
  if (e >= total) return;
  const long b = e / per_ex;
  const int k = (int)(e % bins);
  const float2 s = S[e];
  const float m = sqrtf(fmaf(s.x, s.x, s.y * s.y));
  float2 g = make_float2(0.f, 0.f);
  if (m > 0.f) {
    const float c = upstream[0] * coef[b] * (m - mag_t[e]) / m * ((k == 0 || k == bins - 1) ? scale_n : 0.5f * scale_n);
    g.x = c * s.x; g.y = c * s.y;
  }
  G[e] = g;
}

// dx[b, t] (+)= sum_f dframes[b, f, t - f*hop] * hann(t - f*hop), frames with 0 <= t - f*hop < win
__global__ void __launch_bounds__(256) stft_frames_bwd_kernel(const float* __restrict__ dframes, float* __restrict__ dx, int T, int F,
                                                              int n_fft, int hop, int win, int accumulate, long total) {
  const long e = (long)blockIdx.x * 256 + threadIdx.x;
  if (e >= total) return;
  const int t = (int)(e % T);
  const long b = e / T;
  int f_hi = t / hop;
  if (f_hi > F - 1) f_hi = F - 1;
  int f_lo = t - win + 1 <= 0 ? 0 : (t - win + hop) / hop;  // ceil((t - win + 1) / hop)
  float acc = 0.f;
  for (int f = f_lo; f <= f_hi; ++f) {
    const int n = t - f * hop;
    acc = fmaf(dframes[(b * F + f) * n_fft + n], hann(n, win), acc);
  }
  dx[e] = accumulate ? dx[e] + acc : acc;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_stft_frames(const float* x, int64_t B, int32_t T, int32_t n_fft, int32_t hop, int32_t win, float* frames, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(n_fft >= win && win >= 1 && hop >= 1 && (n_fft & 3) == 0 && T >= win, "vqb_stft_frames: bad STFT geometry");
  const int F = 1 + (T - win) / hop;
  const long total = B * F * (n_fft >> 2);
  if (total == 0) return VQB_OK;
  VQB_REQUIRE(x && frames, "vqb_stft_frames: NULL pointer");
  stft_frames_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, frames, T, F, n_fft, hop, win, total);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

size_t vqb_spec_workspace_bytes(int64_t B, int64_t per_example) { return (size_t)(B * cdiv(per_example, SP_CHUNK) + 16) * sizeof(float); }

/* S [B, per_example] complex64 (interleaved); mag (may be NULL for the diff form) ; sums [B] */
int vqb_spec_mag(const float* S, int64_t B, int64_t per_example, float* mag, float* sums, void* workspace, size_t workspace_bytes,
                 void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(S && mag && sums, "vqb_spec_mag: NULL pointer");
  if (!workspace || workspace_bytes < vqb_spec_workspace_bytes(B, per_example)) return set_err(VQB_ERR_WORKSPACE, "vqb_spec_mag: workspace too small");
  const int chunks = cdiv(per_example, SP_CHUNK);
  if (B == 0 || chunks == 0) return VQB_OK;
  spec_mag_kernel<<<dim3(chunks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>((const float2*)S, mag, per_example, (float*)workspace);
  VQB_LAUNCH_CHECK();
  spec_sum_kernel<<<cdiv(B, 8), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, chunks, (int)B, sums);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_spec_diff(const float* S, const float* mag_t, int64_t B, int64_t per_example, float* sums, void* workspace,
                  size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(S && mag_t && sums, "vqb_spec_diff: NULL pointer");
  if (!workspace || workspace_bytes < vqb_spec_workspace_bytes(B, per_example)) return set_err(VQB_ERR_WORKSPACE, "vqb_spec_diff: workspace too small");
  const int chunks = cdiv(per_example, SP_CHUNK);
  if (B == 0 || chunks == 0) return VQB_OK;
  spec_diff_kernel<<<dim3(chunks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>((const float2*)S, mag_t, per_example, (float*)workspace);
  VQB_LAUNCH_CHECK();
  spec_sum_kernel<<<cdiv(B, 8), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, chunks, (int)B, sums);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

/* dsum, tsum [nscales, B]; loss [1]; coef [nscales, B] (may be NULL) */
int vqb_spec_loss(const float* dsum, const float* tsum, int32_t nscales, int32_t B, float* loss, float* coef, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(dsum && tsum && loss && nscales >= 1 && B >= 1, "vqb_spec_loss: bad arguments");
  spec_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dsum, tsum, nscales, B, loss, coef);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

/* G [B, F, bins] complex64 = the spectrum to feed irfft(n = n_fft) so that its output is d loss / d frames */
int vqb_spec_grad(const float* S, const float* mag_t, const float* coef, const float* upstream, int64_t B, int64_t per_example,
                  int32_t bins, int32_t n_fft, float* G, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(S && mag_t && coef && upstream && G && bins >= 2, "vqb_spec_grad: bad arguments");
  const long total = B * per_example;
  if (total == 0) return VQB_OK;
  // scale 1: the caller's inverse FFT is the UNNORMALISED one (irfft(..., norm="forward")), which saves the library's 1/n pass
  // over the frames; n_fft is a power of two in every STFT scale of the reference, so the values are bit-identical to scaling
  // by n here and by 1/n there
  spec_grad_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)S, mag_t, coef, upstream, per_example, bins,
                                                                        1.0f, (float2*)G, total);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_stft_frames_bwd(const float* dframes, int64_t B, int32_t T, int32_t n_fft, int32_t hop, int32_t win, int32_t accumulate,
                        float* dx, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(n_fft >= win && win >= 1 && hop >= 1 && T >= win, "vqb_stft_frames_bwd: bad STFT geometry");
  const int F = 1 + (T - win) / hop;
  const long total = B * T;
  if (total == 0) return VQB_OK;
  VQB_REQUIRE(dframes && dx, "vqb_stft_frames_bwd: NULL pointer");
  stft_frames_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(dframes, dx, T, F, n_fft, hop, win, accumulate, total);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

}  // extern "C"
