// conv_fp32.cu — exact-fp32 (CUDA-core FFMA) Conv1D / Conv1DTranspose forward, data-gradient and
// weight-gradient kernels behind vqb_conv1d_* / vqb_conv1d_transpose_* (include/vqb.h).
//
// Every one of those ops is a "tap-gather contraction"
//     out[b, t*out_step + out_off, o] = sum_n sum_i act(in[b, t*in_step + off[n], i]) * W_{wj[n]}[i, o]
// over a small host-built tap table:
//   Conv1D forward            (encdec.py:33,38,60,148; resnet.py:13,17)  in_step = stride, off = j*dil - padL
//   Conv1D data gradient      one launch per output phase r < stride: off = (r + padL - j*dil)/stride
//   Conv1DTranspose forward   (encdec.py:67-68) same phase form with dil = 1, padL = (k - stride)/2
//   Conv1DTranspose data grad Conv1D-forward form over dy
// so one kernel (tgc_kernel) serves all four; tgc_narrow_kernel covers C_out <= 4 (the final 64->1 conv),
// wgrad_kernel the weight gradients (split over time chunks, reduced in a fixed order => deterministic).
#include <stdlib.h>

#include "common.cuh"

namespace vqb {

constexpr int MAX_TAPS = 16;
struct TapTable {
  int ntaps;
  int wj[MAX_TAPS];
  int off[MAX_TAPS];
};

struct TgcParams {
  const float* in;
  const float* w;
  const float* bias;  // [COUT] or null
  const float* res;   // added before masking (forward residual) or null
  const float* mask;  // out *= (mask > 0) or null     (ReLU backward)
  const float* add;   // out += add after masking or null (skip-path gradient)
  float* out;
  int B, L_in, CIN, COUT, L_out;  // L_out: rows per batch item of out/res/mask/add
  int Lt, in_step, out_step, out_off;
  int relu_in;
  int w_sj, w_si, w_so;  // W_j[i,o] = w[j*w_sj + i*w_si + o*w_so]
  int ci_ch, rows, minoff;
  TapTable taps;
};

constexpr int TGC_TT = 128;

// 256 threads; tile = 128 time positions x 32 output channels; thread = 4 positions x 4 channels.
__global__ void __launch_bounds__(256) tgc_kernel(const TgcParams p) {
  pdl_launch_dependents();
  pdl_wait();  // launched through launch_pdl (common.cuh): nothing is read or written above this line
  extern __shared__ __align__(16) float smem[];
  float* in_s = smem;                            // [rows][ci_ch]
  float* w_s = smem + (size_t)p.rows * p.ci_ch;  // [ntaps][ci_ch][32]
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int t0 = blockIdx.x * TGC_TT, b = blockIdx.y, co0 = blockIdx.z * 32;
  const int cch = p.ci_ch;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

  const float* inb = p.in + (size_t)b * p.L_in * p.CIN;
  const long g0 = (long)t0 * p.in_step + p.minoff;
  const int ntaps = p.taps.ntaps;

  for (int c0 = 0; c0 < p.CIN && ntaps > 0; c0 += cch) {
    if ((p.CIN & 3) == 0) {
      const int v4 = cch >> 2;
      for (int e = tid; e < p.rows * v4; e += 256) {
        const int r = e / v4, c = (e - r * v4) * 4;
        const long g = g0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g >= 0 && g < p.L_in && c0 + c < p.CIN) v = *(const float4*)(inb + g * p.CIN + c0 + c);
        if (p.relu_in) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        *(float4*)(in_s + r * cch + c) = v;
      }
    } else {
      for (int e = tid; e < p.rows * cch; e += 256) {
        const int r = e / cch, c = e - r * cch;
        const long g = g0 + r;
        float v = 0.f;
        if (g >= 0 && g < p.L_in && c0 + c < p.CIN) v = inb[g * p.CIN + c0 + c];
        if (p.relu_in) v = fmaxf(v, 0.f);
        in_s[r * cch + c] = v;
      }
    }
    for (int e = tid; e < ntaps * cch * 32; e += 256) {
      const int o = e & 31, c = (e >> 5) % cch, n = (e >> 5) / cch;
      const int ci = c0 + c, co = co0 + o;
      float v = 0.f;
      if (ci < p.CIN && co < p.COUT) v = p.w[(size_t)p.taps.wj[n] * p.w_sj + (size_t)ci * p.w_si + (size_t)co * p.w_so];
      w_s[e] = v;
    }
    __syncthreads();
    const int rstride = 32 * p.in_step * cch;
    for (int n = 0; n < ntaps; ++n) {
      const float* a0 = in_s + (size_t)(ty * p.in_step + p.taps.off[n] - p.minoff) * cch;
      const float* wn = w_s + n * cch * 32 + tx * 4;
      for (int c = 0; c < cch; c += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *(const float4*)(a0 + i * rstride + c);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 wv = *(const float4*)(wn + (c + cc) * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float av = cc == 0 ? a[i].x : cc == 1 ? a[i].y : cc == 2 ? a[i].z : a[i].w;
            acc[i][0] = fmaf(av, wv.x, acc[i][0]);
            acc[i][1] = fmaf(av, wv.y, acc[i][1]);
            acc[i][2] = fmaf(av, wv.z, acc[i][2]);
            acc[i][3] = fmaf(av, wv.w, acc[i][3]);
          }
        }
      }
    }
    __syncthreads();
  }

  const int co = co0 + tx * 4;
  if (co >= p.COUT) return;
  const bool vec = ((p.COUT & 3) == 0);
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (co + c < p.COUT) bv[c] = p.bias[co + c];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty + 32 * i;
    if (t >= p.Lt) continue;
    const size_t row = (size_t)b * p.L_out + (size_t)t * p.out_step + p.out_off;
    const size_t base = row * p.COUT + co;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = acc[i][c] + bv[c];
    if (vec) {
      if (p.res) { const float4 r = *(const float4*)(p.res + base); v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w; }
      if (p.mask) {
        const float4 m = *(const float4*)(p.mask + base);
        v[0] = m.x > 0.f ? v[0] : 0.f; v[1] = m.y > 0.f ? v[1] : 0.f;
        v[2] = m.z > 0.f ? v[2] : 0.f; v[3] = m.w > 0.f ? v[3] : 0.f;
      }
      if (p.add) { const float4 r = *(const float4*)(p.add + base); v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w; }
      *(float4*)(p.out + base) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (co + c >= p.COUT) break;
        float x = v[c];
        if (p.res) x += p.res[base + c];
        if (p.mask) x = p.mask[base + c] > 0.f ? x : 0.f;
        if (p.add) x += p.add[base + c];
        p.out[base + c] = x;
      }
    }
  }
}

// C_out <= 4 (e.g. the final Conv1D(1, 3), encdec.py:148): one thread per output position, weights in smem.
__global__ void __launch_bounds__(256) tgc_narrow_kernel(const TgcParams p) {
  pdl_launch_dependents();
  pdl_wait();  // launched through launch_pdl (common.cuh): nothing is read or written above this line
  extern __shared__ __align__(16) float w_s[];  // [ntaps][CIN][COUT]
  const int ntaps = p.taps.ntaps;
  for (int e = threadIdx.x; e < ntaps * p.CIN * p.COUT; e += blockDim.x) {
    const int o = e % p.COUT, ci = (e / p.COUT) % p.CIN, n = e / (p.COUT * p.CIN);
    w_s[e] = p.w[(size_t)p.taps.wj[n] * p.w_sj + (size_t)ci * p.w_si + (size_t)o * p.w_so];
  }
  __syncthreads();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)p.B * p.Lt) return;
  const int b = (int)(idx / p.Lt), t = (int)(idx - (long)b * p.Lt);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* inb = p.in + (size_t)b * p.L_in * p.CIN;
  for (int n = 0; n < ntaps; ++n) {
    const long g = (long)t * p.in_step + p.taps.off[n];
    if (g < 0 || g >= p.L_in) continue;
    const float* row = inb + g * p.CIN;
    const float* wn = w_s + (size_t)n * p.CIN * p.COUT;
    if ((p.CIN & 3) == 0) {
      for (int ci = 0; ci < p.CIN; ci += 4) {
        float4 a = *(const float4*)(row + ci);
        if (p.relu_in) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        for (int o = 0; o < p.COUT; ++o) {
          acc[o] = fmaf(a.x, wn[(ci + 0) * p.COUT + o], acc[o]);
          acc[o] = fmaf(a.y, wn[(ci + 1) * p.COUT + o], acc[o]);
          acc[o] = fmaf(a.z, wn[(ci + 2) * p.COUT + o], acc[o]);
          acc[o] = fmaf(a.w, wn[(ci + 3) * p.COUT + o], acc[o]);
        }
      }
    } else {
      for (int ci = 0; ci < p.CIN; ++ci) {
        float a = row[ci];
        if (p.relu_in) a = fmaxf(a, 0.f);
        for (int o = 0; o < p.COUT; ++o) acc[o] = fmaf(a, wn[ci * p.COUT + o], acc[o]);
      }
    }
  }
  const size_t base = ((size_t)b * p.L_out + (size_t)t * p.out_step + p.out_off) * p.COUT;
  for (int o = 0; o < p.COUT; ++o) {
    float x = acc[o] + (p.bias ? p.bias[o] : 0.f);
    if (p.res) x += p.res[base + o];
    if (p.mask) x = p.mask[base + o] > 0.f ? x : 0.f;
    if (p.add) x += p.add[base + o];
    p.out[base + o] = x;
  }
}

static int launch_tgc(TgcParams& p, cudaStream_t st) {
  int minoff = 0, maxoff = 0;
  for (int n = 0; n < p.taps.ntaps; ++n) {
    if (n == 0 || p.taps.off[n] < minoff) minoff = p.taps.off[n];
    if (n == 0 || p.taps.off[n] > maxoff) maxoff = p.taps.off[n];
  }
  if (p.Lt <= 0 || p.B <= 0) return VQB_OK;
  if (p.COUT <= 4) {
    const size_t smem = (size_t)p.taps.ntaps * p.CIN * p.COUT * sizeof(float);
    VQB_REQUIRE(smem <= 48 * 1024, "narrow conv: weights (%zu B) exceed 48 KB of shared memory", smem);
    const long n = (long)p.B * p.Lt;
    VQB_CUDA(launch_pdl2(tgc_narrow_kernel, dim3(cdiv(n, 256)), dim3(256), (size_t)smem, st, p));
    VQB_LAUNCH_CHECK();
    return VQB_OK;
  }
  p.minoff = minoff;
  p.ci_ch = p.CIN >= 32 ? 32 : ((p.CIN + 3) & ~3);
  p.rows = (TGC_TT - 1) * p.in_step + (maxoff - minoff) + 1;
  size_t smem = ((size_t)p.rows * p.ci_ch + (size_t)p.taps.ntaps * p.ci_ch * 32) * sizeof(float);
  while (smem > 200 * 1024 && p.ci_ch > 4) {  // very large dilation: narrow the channel chunk
    p.ci_ch >>= 1;
    smem = ((size_t)p.rows * p.ci_ch + (size_t)p.taps.ntaps * p.ci_ch * 32) * sizeof(float);
  }
  VQB_REQUIRE(smem <= 200 * 1024, "conv: receptive field of one tile (%d rows) does not fit shared memory", p.rows);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    VQB_CUDA(cudaFuncSetAttribute(tgc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  dim3 grid(cdiv(p.Lt, TGC_TT), p.B, cdiv(p.COUT, 32));
  VQB_CUDA(launch_pdl2(tgc_kernel, dim3(grid), dim3(256), (size_t)smem, st, p));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient: dW[j][cg][co] = sum_{b,t} act(ga[b, t*g_step + off[j], cg]) * ot[b, t, co]
// (Conv1D: ga = x, ot = dy;  Conv1DTranspose: ga = dy, ot = x  =>  both give the Keras kernel layout directly.)
// One CTA = one batch item x WG_TCH time positions x up to 4 taps x a 32x32 channel tile; the two operand tiles
// are staged through shared memory in sub-chunks of WG_SUB positions (coalesced float4 loads, ReLU fused into the
// load), each thread owns NT x 4 x 2 accumulators.  Per-CTA partial sums go to a workspace and are reduced in a fixed
// order (deterministic, no atomics).  When `bias_partial` is set the column sums of `ot` (the Conv1D bias gradient)
// are produced by the same pass.
// ---------------------------------------------------------------------------------------------------------
struct WgParams {
  const float* ga;
  const float* ot;
  float* partial;       // [B*nchunk][ntaps*Cg*Co]
  float* bias_partial;  // [B*nchunk][Co] or null
  int B, Lg, Cg, Lo, Co, Lt, g_step, relu_ga, nchunk, sub, rows, tch;
  TapTable taps;
};

constexpr int WG_TCH = 1024;
constexpr int WG_SUB = 128;

template <int NT>
__global__ void __launch_bounds__(128) wgrad_kernel(const WgParams p) {
  pdl_launch_dependents();
  pdl_wait();  // launched through launch_pdl (common.cuh): nothing is read or written above this line
  extern __shared__ __align__(16) float smem[];
  float* g_s = smem;                            // [rows][32]  gather-side tile
  float* o_s = smem + (size_t)p.rows * 32;      // [sub][32]   other-side tile
  const int tid = threadIdx.x;
  const int ncgt = (p.Cg + 31) >> 5;
  const int tg = blockIdx.y / ncgt, cgt = blockIdx.y - tg * ncgt;  // tap group (4 taps), gather-channel tile
  const int j0 = tg * 4;
  const int cg0 = cgt * 32, co0 = blockIdx.z * 32;
  const int b = blockIdx.x / p.nchunk, ch = blockIdx.x - b * p.nchunk;
  const int tb = ch * p.tch, te = min(tb + p.tch, p.Lt);
  int off[NT], minoff = 1 << 30;
#pragma unroll
  for (int j = 0; j < NT; ++j) { off[j] = p.taps.off[j0 + j]; minoff = min(minoff, off[j]); }
  const float* gab = p.ga + (size_t)b * p.Lg * p.Cg;
  const float* otb = p.ot + (size_t)b * p.Lo * p.Co;
  const bool gvec = (p.Cg & 3) == 0, ovec = (p.Co & 3) == 0;
  const int cq = tid >> 4, c2 = tid & 15;  // 4 gather channels cq*4.., 2 other channels c2*2..
  float acc[NT][4][2];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i][0] = acc[j][i][1] = 0.f;
  float bs0 = 0.f, bs1 = 0.f;
  const bool do_bias = p.bias_partial != nullptr && blockIdx.y == 0 && cq == 0;

  for (int t0 = tb; t0 < te; t0 += p.sub) {
    const int nt = min(p.sub, te - t0);
    const long gr0 = (long)t0 * p.g_step + minoff;
    const int nrows = (nt - 1) * p.g_step + (p.rows - (p.sub - 1) * p.g_step);
    __syncthreads();
    for (int e = tid; e < nrows * 8; e += 128) {
      const int r = e >> 3, c = (e & 7) * 4;
      const long gr = gr0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr >= 0 && gr < p.Lg) {
        const float* src = gab + gr * p.Cg + cg0 + c;
        if (gvec) { if (cg0 + c < p.Cg) v = *(const float4*)src; }
        else {
          if (cg0 + c + 0 < p.Cg) v.x = src[0];
          if (cg0 + c + 1 < p.Cg) v.y = src[1];
          if (cg0 + c + 2 < p.Cg) v.z = src[2];
          if (cg0 + c + 3 < p.Cg) v.w = src[3];
        }
        if (p.relu_ga) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      *(float4*)(g_s + r * 32 + c) = v;
    }
    for (int e = tid; e < nt * 8; e += 128) {
      const int r = e >> 3, c = (e & 7) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = otb + (size_t)(t0 + r) * p.Co + co0 + c;
      if (ovec) { if (co0 + c < p.Co) v = *(const float4*)src; }
      else {
        if (co0 + c + 0 < p.Co) v.x = src[0];
        if (co0 + c + 1 < p.Co) v.y = src[1];
        if (co0 + c + 2 < p.Co) v.z = src[2];
        if (co0 + c + 3 < p.Co) v.w = src[3];
      }
      *(float4*)(o_s + r * 32 + c) = v;
    }
    __syncthreads();
    const float* gp = g_s + cq * 4;
    const float* op = o_s + c2 * 2;
#pragma unroll 2
    for (int t = 0; t < nt; ++t) {
      const float2 g2 = *(const float2*)(op + t * 32);
      if (do_bias) { bs0 += g2.x; bs1 += g2.y; }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float4 a = *(const float4*)(gp + (t * p.g_step + off[j] - minoff) * 32);
        acc[j][0][0] = fmaf(a.x, g2.x, acc[j][0][0]); acc[j][0][1] = fmaf(a.x, g2.y, acc[j][0][1]);
        acc[j][1][0] = fmaf(a.y, g2.x, acc[j][1][0]); acc[j][1][1] = fmaf(a.y, g2.y, acc[j][1][1]);
        acc[j][2][0] = fmaf(a.z, g2.x, acc[j][2][0]); acc[j][2][1] = fmaf(a.z, g2.y, acc[j][2][1]);
        acc[j][3][0] = fmaf(a.w, g2.x, acc[j][3][0]); acc[j][3][1] = fmaf(a.w, g2.y, acc[j][3][1]);
      }
    }
  }
  float* out = p.partial + (size_t)blockIdx.x * p.taps.ntaps * p.Cg * p.Co;
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cg = cg0 + cq * 4 + i, co = co0 + c2 * 2 + c;
        if (cg < p.Cg && co < p.Co) out[((size_t)(j0 + j) * p.Cg + cg) * p.Co + co] = acc[j][i][c];
      }
  if (do_bias) {
    const int co = co0 + c2 * 2;
    if (co < p.Co) p.bias_partial[(size_t)blockIdx.x * p.Co + co] = bs0;
    if (co + 1 < p.Co) p.bias_partial[(size_t)blockIdx.x * p.Co + co + 1] = bs1;
  }
}

// out[e] = sum_{c < nchunk} partial[c*stride + offset + e]; block = 32 elements x 8 chunk lanes, fixed summation tree
__global__ void __launch_bounds__(256) reduce_chunks_kernel(const float* __restrict__ partial, int nchunk, long stride,
                                                            int offset, int n, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int ex = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + ex;
  float s = 0.f;
  if (e < n)
    for (int c = ly; c < nchunk; c += 8) s += partial[(size_t)c * stride + offset + e];
  red[ly][ex] = s;
  __syncthreads();
  if (ly == 0 && e < n) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += red[l][ex];
    out[e] = t;
  }
}

// ---- deferred / batched reductions -------------------------------------------------------------------------
// Between vqb_reduce_begin() and vqb_reduce_flush() the fixed-order reductions requested by the weight-gradient entry
// points are queued (thread-local) instead of launched, and flushed as a few batched kernels: a training step has
// ~480 of them (kernel + bias per convolution), each only a few microseconds of work.
constexpr int RB_ITEMS = 64;
struct ReduceItem {
  const float* partial;
  float* out;
  int nchunk, stride, offset, n, block0;
};
struct ReduceBatch {
  ReduceItem it[RB_ITEMS];
  int n_items;
};

__global__ void __launch_bounds__(256) reduce_batched_kernel(const ReduceBatch b) {
  pdl_launch_dependents();
  pdl_wait();  // launched through launch_pdl (common.cuh): nothing is read or written above this line
  __shared__ float red[8][33];
  int lo = 0, hi = b.n_items - 1;  // last item whose block0 <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.it[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const ReduceItem& t = b.it[lo];
  const int ex = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int e = ((int)blockIdx.x - t.block0) * 32 + ex;
  float s = 0.f;
  if (e < t.n)
    for (int c = ly; c < t.nchunk; c += 8) s += t.partial[(size_t)c * t.stride + t.offset + e];
  red[ly][ex] = s;
  __syncthreads();
  if (ly == 0 && e < t.n) {
    float v = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) v += red[l][ex];
    t.out[e] = v;
  }
}

struct ReduceQueue {
  bool active = false;
  int n = 0, cap = 0;
  ReduceItem* items = nullptr;
};
static thread_local ReduceQueue g_rq;

void reduce_chunks_strided(const float* partial, int nchunk, long stride, int offset, int n, float* out, cudaStream_t st) {
  if (g_rq.active) {
    if (g_rq.n == g_rq.cap) {
      g_rq.cap = g_rq.cap ? 2 * g_rq.cap : 1024;
      g_rq.items = (ReduceItem*)realloc(g_rq.items, sizeof(ReduceItem) * g_rq.cap);
    }
    g_rq.items[g_rq.n++] = ReduceItem{partial, out, nchunk, (int)stride, offset, n, 0};
    uncount_launch();  // the caller's VQB_LAUNCH_CHECK counts a launch that happens, batched, in reduce_flush
    return;
  }
  reduce_chunks_kernel<<<cdiv(n, 32), 256, 0, st>>>(partial, nchunk, stride, offset, n, out);
}

int reduce_flush(cudaStream_t st) {
  int i = 0;
  while (i < g_rq.n) {
    ReduceBatch b;
    int blocks = 0, k = 0;
    for (; k < RB_ITEMS && i + k < g_rq.n; ++k) {
      b.it[k] = g_rq.items[i + k];
      b.it[k].block0 = blocks;
      blocks += cdiv(b.it[k].n, 32);
    }
    b.n_items = k;
    if (blocks > 0) {
      VQB_CUDA(launch_pdl2(reduce_batched_kernel, dim3(blocks), dim3(256), (size_t)0, st, b));
      VQB_LAUNCH_CHECK();
    }
    i += k;
  }
  g_rq.n = 0;
  g_rq.active = false;
  return VQB_OK;
}

void reduce_begin() { g_rq.active = true; g_rq.n = 0; }

// column sums of a [rows, C] matrix (Conv1DTranspose bias gradient): partial[block][c]
constexpr int COLSUM_ROWS = 256;  // rows per CTA: small, so that a [28160, 64] matrix still spreads over ~110 CTAs
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long rows, int C, float* __restrict__ partial) {
  __shared__ float red[256];
  const long r0 = (long)blockIdx.x * COLSUM_ROWS, r1 = min(r0 + (long)COLSUM_ROWS, rows);
  if (C <= 256) {
    const int lanes = 256 / C, nt = lanes * C;
    const int c = threadIdx.x % C, rl = threadIdx.x / C;
    float s = 0.f;
    if (threadIdx.x < nt) {  // four independent chains keep four loads in flight per thread (fixed order: deterministic)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      long r = r0 + rl;
      for (; r + 3 * lanes < r1; r += 4 * lanes) {
        s0 += x[r * C + c]; s1 += x[(r + lanes) * C + c]; s2 += x[(r + 2 * lanes) * C + c]; s3 += x[(r + 3 * lanes) * C + c];
      }
      for (; r < r1; r += lanes) s0 += x[r * C + c];
      s = (s0 + s1) + (s2 + s3);
    }
    red[threadIdx.x] = threadIdx.x < nt ? s : 0.f;
    __syncthreads();
    if (threadIdx.x < C) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += red[l * C + threadIdx.x];
      partial[(size_t)blockIdx.x * C + threadIdx.x] = t;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += 256) {
      float s = 0.f;
      for (long r = r0; r < r1; ++r) s += x[r * C + c];
      partial[(size_t)blockIdx.x * C + c] = s;
    }
  }
}

// time positions per CTA: 1024, halved while the grid would leave most of the 148 SMs idle (short sequences)
static int pick_tch(int B, int Lt, int ntaps, int Cg, int Co) {
  int tch = WG_TCH;
  const long yz = (long)cdiv(ntaps, 4) * cdiv(Cg, 32) * cdiv(Co, 32);
  while (tch > WG_SUB && (long)B * cdiv(Lt, tch) * yz < 592) tch >>= 1;
  return tch;
}

// workspace: [B*nchunk][ntaps*Cg*Co] weight partials, then the bias partials
static size_t wgrad_ws_floats(int B, int Lt, int ntaps, int Cg, int Co, long bias_rows, int Cb) {
  const size_t nchunk = (size_t)B * cdiv(Lt, pick_tch(B, Lt, ntaps, Cg, Co));
  const size_t nb = (size_t)cdiv(bias_rows, COLSUM_ROWS);
  return nchunk * ntaps * Cg * Co + (nchunk > nb ? nchunk : nb) * Cb + 64;
}

template <int NT>
static int launch_wgrad(const WgParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    static bool set = false;
    if (!set) { VQB_CUDA(cudaFuncSetAttribute(wgrad_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); set = true; }
  }
  VQB_CUDA(launch_pdl2(wgrad_kernel<NT>, dim3(grid), dim3(128), (size_t)smem, st, p));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

// bias_from_ot: the bias gradient is the column sum of `ot` (Conv1D) and is fused; otherwise it is the column sum of
// `bias_src` (Conv1DTranspose: all rows of dy) through colsum_kernel.
static int run_wgrad(WgParams& p, float* dw, bool bias_from_ot, const float* bias_src, long bias_rows, int Cb,
                     float* dbias, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int ntaps = p.taps.ntaps;
  p.tch = pick_tch(p.B, p.Lt, ntaps, p.Cg, p.Co);
  p.nchunk = cdiv(p.Lt, p.tch);
  const size_t need = wgrad_ws_floats(p.B, p.Lt, ntaps, p.Cg, p.Co, bias_rows, Cb) * sizeof(float);
  if (ws_bytes < need || !ws) return set_err(VQB_ERR_WORKSPACE, "wgrad workspace: need %zu bytes, got %zu", need, ws_bytes);
  const int nchunks = p.B * p.nchunk;
  const int n = ntaps * p.Cg * p.Co;
  p.partial = (float*)ws;
  float* bp = p.partial + (size_t)nchunks * n;
  p.bias_partial = (dbias && bias_from_ot) ? bp : nullptr;
  if (nchunks == 0) {
    VQB_CUDA(cudaMemsetAsync(dw, 0, (size_t)n * sizeof(float), st));
    if (dbias) VQB_CUDA(cudaMemsetAsync(dbias, 0, (size_t)Cb * sizeof(float), st));
    return VQB_OK;
  }
  // taps are processed 4 per CTA; pad the table so that every group is full (padding taps repeat the last offset and
  // write to a scratch slot that is never read: handled by giving them weight index ntaps.. < 4*ngroups)
  int span = 0;
  for (int g0 = 0; g0 < ntaps; g0 += 4) {
    int lo = 1 << 30, hi = -(1 << 30);
    for (int j = g0; j < g0 + 4 && j < ntaps; ++j) { lo = p.taps.off[j] < lo ? p.taps.off[j] : lo; hi = p.taps.off[j] > hi ? p.taps.off[j] : hi; }
    span = hi - lo > span ? hi - lo : span;
  }
  p.sub = WG_SUB;
  size_t smem;
  for (;;) {
    p.rows = (p.sub - 1) * p.g_step + span + 1;
    smem = ((size_t)p.rows * 32 + (size_t)p.sub * 32) * sizeof(float);
    if (smem <= 160 * 1024 || p.sub <= 8) break;
    p.sub >>= 1;
  }
  VQB_REQUIRE(smem <= 160 * 1024, "wgrad: tap span %d does not fit shared memory", span);
  const int full = ntaps / 4, rem = ntaps % 4;
  const int ncgt = cdiv(p.Cg, 32);
  if (full > 0) {
    dim3 grid(nchunks, full * ncgt, cdiv(p.Co, 32));
    int rc = launch_wgrad<4>(p, grid, smem, st);
    if (rc) return rc;
  }
  if (rem > 0) {
    WgParams q = p;  // remaining taps: shift them to the front of a private table
    for (int j = 0; j < rem; ++j) { q.taps.off[j] = p.taps.off[full * 4 + j]; }
    q.partial = p.partial + (size_t)full * 4 * p.Cg * p.Co;  // tap index offset inside each chunk's block
    q.bias_partial = full > 0 ? nullptr : p.bias_partial;
    dim3 grid(nchunks, ncgt, cdiv(p.Co, 32));
    int rc = rem == 1 ? launch_wgrad<1>(q, grid, smem, st) : rem == 2 ? launch_wgrad<2>(q, grid, smem, st)
                                                                      : launch_wgrad<3>(q, grid, smem, st);
    if (rc) return rc;
  }
  reduce_chunks_strided(p.partial, nchunks, n, 0, n, dw, st);
  VQB_LAUNCH_CHECK();
  if (dbias) {
    if (bias_from_ot) {
      reduce_chunks_strided(bp, nchunks, Cb, 0, Cb, dbias, st);
      VQB_LAUNCH_CHECK();
    } else {
      const int nb = cdiv(bias_rows, COLSUM_ROWS);
      colsum_kernel<<<nb, 256, 0, st>>>(bias_src, bias_rows, Cb, bp);
      VQB_LAUNCH_CHECK();
      reduce_chunks_strided(bp, nb, Cb, 0, Cb, dbias, st);
      VQB_LAUNCH_CHECK();
    }
  }
  return VQB_OK;
}

static inline int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
static inline int floormod(int a, int b) { return a - floordiv(a, b) * b; }

static int check_desc(const vqb_conv_desc* d, bool transpose) {
  VQB_REQUIRE(d != nullptr, "conv desc is NULL");
  VQB_REQUIRE(d->B >= 0 && d->L >= 0 && d->C_in > 0 && d->C_out > 0, "conv desc: bad shape B=%d L=%d Cin=%d Cout=%d",
              d->B, d->L, d->C_in, d->C_out);
  VQB_REQUIRE(d->k >= 1 && d->k <= MAX_TAPS, "conv desc: k=%d outside [1,%d]", d->k, MAX_TAPS);
  VQB_REQUIRE(d->stride >= 1 && d->dilation >= 1, "conv desc: stride=%d dilation=%d", d->stride, d->dilation);
  if (transpose) VQB_REQUIRE(d->dilation == 1, "Conv1DTranspose: dilation must be 1");
  return VQB_OK;
}

static void same_pad(int L, int k, int s, int dil, int* out, int* left) {
  *out = (L + s - 1) / s;
  int pad = (*out - 1) * s + (k - 1) * dil + 1 - L;
  if (pad < 0) pad = 0;
  *left = pad / 2;
}

// the "phase" form shared by Conv1D data-gradient and Conv1DTranspose forward
static int run_phases(TgcParams base, int k, int s, int dil, int padL, int L_full, cudaStream_t st) {
  for (int r = 0; r < s; ++r) {
    TgcParams p = base;
    p.taps.ntaps = 0;
    for (int j = 0; j < k; ++j) {
      const int num = r + padL - j * dil;
      if (floormod(num, s) != 0) continue;
      p.taps.wj[p.taps.ntaps] = j;
      p.taps.off[p.taps.ntaps] = floordiv(num, s);
      ++p.taps.ntaps;
    }
    p.in_step = 1;
    p.out_step = s;
    p.out_off = r;
    p.Lt = L_full > r ? (L_full - r + s - 1) / s : 0;
    int rc = launch_tgc(p, st);
    if (rc != VQB_OK) return rc;
  }
  return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// C_in = 1 (the first encoder convolution, encdec.py:33 on the raw waveform): a k-tap FIR into C_out <= 32 channels.
// Pure bandwidth: the forward pass writes y once; the weight/bias gradient reads dy once (lane = output channel, the
// k waveform samples of a row are fetched by k lanes and broadcast), per-CTA partials, fixed-order reduction.
// ---------------------------------------------------------------------------------------------------------
constexpr int IN1_MAXK = 8;
__global__ void __launch_bounds__(256) conv_in1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y, int L, int Lo,
                                                           int Co, int k, int stride, int padL, long total) {
  __shared__ float ws[IN1_MAXK * 32 + 32];
  for (int e = threadIdx.x; e < k * Co; e += 256) ws[e] = w[e];
  for (int e = threadIdx.x; e < Co; e += 256) ws[IN1_MAXK * 32 + e] = bias ? bias[e] : 0.f;
  __syncthreads();
  const long e = (long)blockIdx.x * 256 + threadIdx.x;  // one thread per (row, 4 channels)
  if (e >= total) return;
  const int c4n = Co >> 2;
  const int c = (int)(e % c4n) * 4;
  const long row = e / c4n;
  const long b = row / Lo;
  const int t = (int)(row - b * Lo);
  const float* xb = x + b * L;
  float4 acc = *reinterpret_cast<const float4*>(ws + IN1_MAXK * 32 + c);
  for (int j = 0; j < k; ++j) {
    const int g = t * stride + j - padL;
    if (g >= 0 && g < L) {
      const float xv = xb[g];
      const float4 wv = *reinterpret_cast<const float4*>(ws + j * Co + c);
      acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
    }
  }
  *reinterpret_cast<float4*>(y + row * Co + c) = acc;
}

constexpr int IN1_PART = (IN1_MAXK + 1) * 32;  // per-CTA partial: k tap rows + the bias row, 32 channels each
template <int KT>  // taps handled (>= k): 4 for the encoders' Conv1D(32, 4, strides=2), else IN1_MAXK
__global__ void __launch_bounds__(256) conv_in1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             float* __restrict__ partial, long rows, int L, int Lo, int Co, int k,
                                                             int stride, int padL, long rpw) {
  // A warp walks its rows four at a time: lane = (row in group, channel quad), so one warp instruction fetches 4 x 128 B of dy,
  // and four groups are in flight per warp (unroll): the kernel is a pure stream over dy, its speed is the number of bytes in
  // flight per SM (the one-row-per-instruction form of round 1 had 16 KB in flight per SM and ran at a fifth of the HBM rate).
  __shared__ float red[8][IN1_PART];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rs = lane >> 3, q = lane & 7;
  float acc[KT + 1][4];
#pragma unroll
  for (int j = 0; j <= KT; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
  const long r0 = ((long)blockIdx.x * 8 + warp) * rpw;
  const long r1 = r0 + rpw < rows ? r0 + rpw : rows;
  const bool qok = q * 4 < Co;
  long b = (r0 + rs) / Lo;           // one 64-bit division per thread, not per row
  int t = (int)(r0 + rs - b * Lo);
#pragma unroll 4
  for (long row = r0 + rs; row < r1; row += 4) {
    const float4 d = qok ? *reinterpret_cast<const float4*>(dy + row * Co + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xb = x + b * L;
    const int g0 = t * stride - padL;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const int g = g0 + j;
      const float xv = (j < k && g >= 0 && g < L) ? xb[g] : 0.f;
      acc[j][0] = fmaf(xv, d.x, acc[j][0]); acc[j][1] = fmaf(xv, d.y, acc[j][1]);
      acc[j][2] = fmaf(xv, d.z, acc[j][2]); acc[j][3] = fmaf(xv, d.w, acc[j][3]);
    }
    acc[KT][0] += d.x; acc[KT][1] += d.y; acc[KT][2] += d.z; acc[KT][3] += d.w;
    t += 4;
    while (t >= Lo) { t -= Lo; ++b; }
  }
  // the four row groups of a warp (lanes q, q + 8, q + 16, q + 24), fixed order
  for (int e = lane; e < IN1_PART; e += 32) red[warp][e] = 0.f;  // tap rows KT .. IN1_MAXK - 1 stay zero
  __syncwarp();
#pragma unroll
  for (int j = 0; j <= KT; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = acc[j][c];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (rs == 0) red[warp][(j < KT ? j : IN1_MAXK) * 32 + q * 4 + c] = v;  // the bias row keeps its place behind IN1_MAXK tap rows
    }
  __syncthreads();
  for (int e = tid; e < IN1_PART; e += 256) {
    float s_ = red[0][e];
#pragma unroll
    for (int w_ = 1; w_ < 8; ++w_) s_ += red[w_][e];
    partial[(size_t)blockIdx.x * IN1_PART + e] = s_;
  }
}

static bool conv_in1_ok(const vqb_conv_desc* d) {
  return d->C_in == 1 && d->C_out <= 32 && (d->C_out & 3) == 0 && d->k <= IN1_MAXK && d->dilation == 1 && !d->relu_in;
}
static int conv_in1_grid(long rows, long* rpw) {
  long warps = cdiv(rows, 64);
  if (warps > 592 * 8) warps = 592 * 8;
  if (warps < 1) warps = 1;
  const int grid = cdiv(warps, 8);
  *rpw = cdiv(rows, (long)grid * 8);
  if (*rpw < 1) *rpw = 1;
  return grid;
}
static size_t conv_in1_ws_floats(const vqb_conv_desc* d) {
  long rpw;
  return (size_t)conv_in1_grid((long)d->B * ((d->L + d->stride - 1) / d->stride), &rpw) * IN1_PART + 64;
}

int conv1d_fwd_fp32(const vqb_conv_desc* d, const float* x, const float* w, const float* bias,
                    const float* residual, float* y, cudaStream_t st) {
  int Lo, padL;
  same_pad(d->L, d->k, d->stride, d->dilation, &Lo, &padL);
  if (conv_in1_ok(d) && !residual) {
    const long total = (long)d->B * Lo * (d->C_out >> 2);
    if (total == 0) return VQB_OK;
    conv_in1_fwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(x, w, bias, y, d->L, Lo, d->C_out, d->k, d->stride, padL, total);
    VQB_LAUNCH_CHECK();
    return VQB_OK;
  }
  TgcParams p{};
  p.in = x; p.w = w; p.bias = bias; p.res = residual; p.out = y;
  p.B = d->B; p.L_in = d->L; p.CIN = d->C_in; p.COUT = d->C_out; p.L_out = Lo;
  p.Lt = Lo; p.in_step = d->stride; p.out_step = 1; p.out_off = 0; p.relu_in = d->relu_in;
  p.w_sj = d->C_in * d->C_out; p.w_si = d->C_out; p.w_so = 1;
  p.taps.ntaps = d->k;
  for (int j = 0; j < d->k; ++j) { p.taps.wj[j] = j; p.taps.off[j] = j * d->dilation - padL; }
  return launch_tgc(p, st);
}

int conv1d_dgrad_fp32(const vqb_conv_desc* d, const float* dy, const float* w, const float* x,
                      const float* dx_add, float* dx, cudaStream_t st) {
  int Lo, padL;
  same_pad(d->L, d->k, d->stride, d->dilation, &Lo, &padL);
  TgcParams p{};
  p.in = dy; p.w = w; p.mask = d->relu_in ? x : nullptr; p.add = dx_add; p.out = dx;
  p.B = d->B; p.L_in = Lo; p.CIN = d->C_out; p.COUT = d->C_in; p.L_out = d->L;
  p.w_sj = d->C_in * d->C_out; p.w_si = 1; p.w_so = d->C_out;  // W_j[i=co][o=ci] = w[j][ci][co]
  return run_phases(p, d->k, d->stride, d->dilation, padL, d->L, st);
}

int conv1d_transpose_fwd_fp32(const vqb_conv_desc* d, const float* x, const float* w, const float* bias,
                              float* y, cudaStream_t st) {
  const int padL = (d->k > d->stride ? d->k - d->stride : 0) / 2;
  TgcParams p{};
  p.in = x; p.w = w; p.bias = bias; p.out = y;
  p.B = d->B; p.L_in = d->L; p.CIN = d->C_in; p.COUT = d->C_out; p.L_out = d->L * d->stride;
  p.relu_in = d->relu_in;
  p.w_sj = d->C_out * d->C_in; p.w_si = 1; p.w_so = d->C_in;  // W_j[i=ci][o=co] = w[j][co][ci]
  return run_phases(p, d->k, d->stride, 1, padL, d->L * d->stride, st);
}

int conv1d_transpose_dgrad_fp32(const vqb_conv_desc* d, const float* dy, const float* w, float* dx,
                                cudaStream_t st) {
  const int padL = (d->k > d->stride ? d->k - d->stride : 0) / 2;
  TgcParams p{};
  p.in = dy; p.w = w; p.out = dx;
  p.B = d->B; p.L_in = d->L * d->stride; p.CIN = d->C_out; p.COUT = d->C_in; p.L_out = d->L;
  p.Lt = d->L; p.in_step = d->stride; p.out_step = 1; p.out_off = 0;
  p.w_sj = d->C_out * d->C_in; p.w_si = d->C_in; p.w_so = 1;  // W_j[i=co][o=ci] = w[j][co][ci]
  p.taps.ntaps = d->k;
  for (int j = 0; j < d->k; ++j) { p.taps.wj[j] = j; p.taps.off[j] = j - padL; }
  return launch_tgc(p, st);
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_reduce_begin(void) {
  reduce_begin();
  return VQB_OK;
}

int vqb_reduce_flush(void* stream) {
  VQB_ARCH();
  return reduce_flush((cudaStream_t)stream);
}

int vqb_conv1d_fwd(const vqb_conv_desc* d, const float* x, const float* w, const float* bias,
                   const float* residual, float* y, void* stream) {
  VQB_ARCH();
  int rc = check_desc(d, false);
  if (rc) return rc;
  VQB_REQUIRE(x && w && y, "vqb_conv1d_fwd: NULL pointer");
  if (d->precision != VQB_PREC_FP32) {
    if (conv3_tc_supported(d) && !residual) return conv3_fwd_tc(d, x, w, bias, y, (cudaStream_t)stream);
    VQB_REQUIRE(conv_tc_supported(d) && !residual, "vqb_conv1d_fwd: no tensor-core kernel for this shape (k=%d stride=%d %d->%d); use VQB_PREC_FP32",
                d->k, d->stride, d->C_in, d->C_out);
    return conv1d_fwd_tc(d, x, w, bias, y, (cudaStream_t)stream);
  }
  return conv1d_fwd_fp32(d, x, w, bias, residual, y, (cudaStream_t)stream);
}

int vqb_conv1d_dgrad(const vqb_conv_desc* d, const float* dy, const float* w, const float* x,
                     const float* dx_add, float* dx, void* stream) {
  VQB_ARCH();
  int rc = check_desc(d, false);
  if (rc) return rc;
  VQB_REQUIRE(dy && w && dx, "vqb_conv1d_dgrad: NULL pointer");
  VQB_REQUIRE(!d->relu_in || x, "vqb_conv1d_dgrad: relu_in needs x for the ReLU mask");
  if (d->precision != VQB_PREC_FP32) {
    if (conv3_tc_supported(d) && !dx_add) return conv3_dgrad_tc(d, dy, w, dx, (cudaStream_t)stream);
    VQB_REQUIRE(conv_tc_supported(d) && !dx_add, "vqb_conv1d_dgrad: no tensor-core kernel for this shape (k=%d stride=%d %d->%d); use VQB_PREC_FP32",
                d->k, d->stride, d->C_in, d->C_out);
    return conv1d_dgrad_tc(d, dy, w, dx, (cudaStream_t)stream);
  }
  return conv1d_dgrad_fp32(d, dy, w, x, dx_add, dx, (cudaStream_t)stream);
}

int vqb_conv1d_supports(const vqb_conv_desc* d, int op) {
  if (!d || d->k < 1 || d->k > MAX_TAPS) return 0;
  if (d->precision == VQB_PREC_FP32) return 1;
  if (op == 2) return wgrad_tc_supported(d) || conv_tc_supported(d) ? 1 : 0;
  return conv_tc_supported(d) || conv3_tc_supported(d) ? 1 : 0;
}

int vqb_conv1d_transpose_supports(const vqb_conv_desc* d, int op) {
  if (!d || d->k < 1 || d->k > MAX_TAPS) return 0;
  if (d->precision == VQB_PREC_FP32) return 1;
  return conv_tc_supported(d) ? 1 : 0;
}

size_t vqb_conv1d_wgrad_workspace_bytes(const vqb_conv_desc* d) {
  if (!d || d->k < 1 || d->k > MAX_TAPS) return 0;
  if (d->precision != VQB_PREC_FP32 && wgrad_tc_supported(d)) return wgrad_tc_workspace_bytes(d);
  if (d->precision != VQB_PREC_FP32 && conv_tc_supported(d)) return wgrad4_tc_workspace_bytes(d->B, (d->L + 1) / 2);
  int Lo, padL;
  same_pad(d->L, d->k, d->stride, d->dilation, &Lo, &padL);
  size_t n = wgrad_ws_floats(d->B, Lo, d->k, d->C_in, d->C_out, (long)d->B * Lo, d->C_out);
  if (conv_in1_ok(d) && conv_in1_ws_floats(d) > n) n = conv_in1_ws_floats(d);
  return n * sizeof(float);
}

int vqb_conv1d_wgrad(const vqb_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                     void* workspace, size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  int rc = check_desc(d, false);
  if (rc) return rc;
  VQB_REQUIRE(x && dy && dw, "vqb_conv1d_wgrad: NULL pointer");
  if (d->precision != VQB_PREC_FP32) {
    VQB_REQUIRE(wgrad_tc_supported(d) || conv_tc_supported(d),
                "vqb_conv1d_wgrad: no tensor-core kernel for this shape (k=%d stride=%d %d->%d dil=%d); use VQB_PREC_FP32",
                d->k, d->stride, d->C_in, d->C_out, d->dilation);
    if (d->B == 0 || d->L == 0) {
      VQB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * d->k * 32 * 32, (cudaStream_t)stream));
      if (dbias) VQB_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * 32, (cudaStream_t)stream));
      return VQB_OK;
    }
    if (conv_tc_supported(d))
      return wgrad4_tc(d->precision, x, d->L, dy, (d->L + 1) / 2, d->B, dw, dbias, false, workspace, workspace_bytes,
                       (cudaStream_t)stream);
    return conv1d_wgrad_tc(d, x, dy, dw, dbias, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  int Lo, padL;
  same_pad(d->L, d->k, d->stride, d->dilation, &Lo, &padL);
  if (conv_in1_ok(d) && (long)d->B * Lo > 0) {
    const size_t need = conv_in1_ws_floats(d) * sizeof(float);
    if (!workspace || workspace_bytes < need) return set_err(VQB_ERR_WORKSPACE, "wgrad workspace: need %zu bytes, got %zu", need, workspace_bytes);
    long rpw;
    const long rows = (long)d->B * Lo;
    const int grid = conv_in1_grid(rows, &rpw);
    float* partial = (float*)workspace;
    if (d->k <= 4) conv_in1_wgrad_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(x, dy, partial, rows, d->L, Lo, d->C_out, d->k, d->stride, padL, rpw);
    else conv_in1_wgrad_kernel<IN1_MAXK><<<grid, 256, 0, (cudaStream_t)stream>>>(x, dy, partial, rows, d->L, Lo, d->C_out, d->k, d->stride, padL, rpw);
    VQB_LAUNCH_CHECK();
    for (int j = 0; j < d->k; ++j) {  // dw [k, 1, C_out]: one fixed-order reduction per tap row (32-float rows in the partials)
      reduce_chunks_strided(partial, grid, IN1_PART, j * 32, d->C_out, dw + (size_t)j * d->C_out, (cudaStream_t)stream);
      VQB_LAUNCH_CHECK();
    }
    if (dbias) {
      reduce_chunks_strided(partial, grid, IN1_PART, IN1_MAXK * 32, d->C_out, dbias, (cudaStream_t)stream);
      VQB_LAUNCH_CHECK();
    }
    return VQB_OK;
  }
  WgParams p{};
  p.ga = x; p.ot = dy;
  p.B = d->B; p.Lg = d->L; p.Cg = d->C_in; p.Lo = Lo; p.Co = d->C_out; p.Lt = Lo;
  p.g_step = d->stride; p.relu_ga = d->relu_in;
  p.taps.ntaps = d->k;
  for (int j = 0; j < d->k; ++j) { p.taps.wj[j] = j; p.taps.off[j] = j * d->dilation - padL; }
  return run_wgrad(p, dw, true, dy, (long)d->B * Lo, d->C_out, dbias, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vqb_conv1d_transpose_fwd(const vqb_conv_desc* d, const float* x, const float* w, const float* bias,
                             float* y, void* stream) {
  VQB_ARCH();
  int rc = check_desc(d, true);
  if (rc) return rc;
  VQB_REQUIRE(x && w && y, "vqb_conv1d_transpose_fwd: NULL pointer");
  if (d->precision != VQB_PREC_FP32) {
    VQB_REQUIRE(conv_tc_supported(d), "vqb_conv1d_transpose_fwd: no tensor-core kernel for this shape; use VQB_PREC_FP32");
    return conv1d_transpose_fwd_tc(d, x, w, bias, y, (cudaStream_t)stream);
  }
  return conv1d_transpose_fwd_fp32(d, x, w, bias, y, (cudaStream_t)stream);
}

int vqb_conv1d_transpose_dgrad(const vqb_conv_desc* d, const float* dy, const float* w, float* dx, void* stream) {
  VQB_ARCH();
  int rc = check_desc(d, true);
  if (rc) return rc;
  VQB_REQUIRE(dy && w && dx, "vqb_conv1d_transpose_dgrad: NULL pointer");
  if (d->precision != VQB_PREC_FP32) {
    VQB_REQUIRE(conv_tc_supported(d), "vqb_conv1d_transpose_dgrad: no tensor-core kernel for this shape; use VQB_PREC_FP32");
    return conv1d_transpose_dgrad_tc(d, dy, w, dx, (cudaStream_t)stream);
  }
  return conv1d_transpose_dgrad_fp32(d, dy, w, dx, (cudaStream_t)stream);
}

size_t vqb_conv1d_transpose_wgrad_workspace_bytes(const vqb_conv_desc* d) {
  if (!d || d->k < 1 || d->k > MAX_TAPS) return 0;
  if (d->precision != VQB_PREC_FP32 && conv_tc_supported(d)) return wgrad4_tc_workspace_bytes(d->B, d->L);
  return wgrad_ws_floats(d->B, d->L, d->k, d->C_out, d->C_in, (long)d->B * d->L * d->stride, d->C_out) * sizeof(float);
}

int vqb_conv1d_transpose_wgrad(const vqb_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                               void* workspace, size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  int rc = check_desc(d, true);
  if (rc) return rc;
  VQB_REQUIRE(x && dy && dw, "vqb_conv1d_transpose_wgrad: NULL pointer");
  if (d->precision != VQB_PREC_FP32) {
    VQB_REQUIRE(conv_tc_supported(d), "vqb_conv1d_transpose_wgrad: no tensor-core kernel for this shape; use VQB_PREC_FP32");
    if (d->B == 0 || d->L == 0) {
      VQB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * d->k * 32 * 32, (cudaStream_t)stream));
      if (dbias) VQB_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * 32, (cudaStream_t)stream));
      return VQB_OK;
    }
    return wgrad4_tc(d->precision, dy, 2 * d->L, x, d->L, d->B, dw, dbias, true, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  const int padL = (d->k > d->stride ? d->k - d->stride : 0) / 2;
  WgParams p{};
  p.ga = dy; p.ot = x;  // dW[j][co][ci] = sum dy[m*s + j - padL][co] * x[m][ci]
  p.B = d->B; p.Lg = d->L * d->stride; p.Cg = d->C_out; p.Lo = d->L; p.Co = d->C_in; p.Lt = d->L;
  p.g_step = d->stride; p.relu_ga = 0;
  p.taps.ntaps = d->k;
  for (int j = 0; j < d->k; ++j) { p.taps.wj[j] = j; p.taps.off[j] = j - padL; }
  return run_wgrad(p, dw, false, dy, (long)d->B * d->L * d->stride, d->C_out, dbias, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

}  // extern "C"
