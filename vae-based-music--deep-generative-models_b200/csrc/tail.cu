// tail.cu — the decoder's last two layers as ONE linear operator (vqb_dec_tail_*).
//
// The reference decoder ends with Conv1DTranspose(64, k=4, s=2) (encdec.py:67-68, last up-sampling layer of the last
// DecoderConvBlock) followed directly by Conv1D(1, 3) (encdec.py:148): no activation in between, so
//     y[2m]   = x[m] Wt1 + x[m-1] Wt3 + bt,   y[2m+1] = x[m+1] Wt0 + x[m] Wt2 + bt          (Wt[k] is [Cmid, Cin])
//     r[t]    = bf + sum_j Wf[j] . y[t+j-1]    (y zero outside [0, 2L): SAME padding)
// is a 3-tap, Cin -> 2-phase convolution with composed weights:  with V[j][k] = Wt[k]^T Wf[j]  (a Cin-vector) and
// c_j = Wf[j] . bt,
//     r[2m]   = K + x[m-1].(V02+V13) + x[m].(V00+V11+V22) + x[m+1].V20       (t = 0:    minus c_0 + x[0].V00)
//     r[2m+1] = K + x[m-1].V03 + x[m].(V01+V12+V23) + x[m+1].(V10+V21)       (t = 2L-1: minus c_2 + x[L-1].V23)
//     K = bf + c_0 + c_1 + c_2.
// The [B, 2L, 64] intermediate (231 MB at B = 32, T = 28160 — the largest tensor of the model) is never formed: the
// forward pass reads x once and writes r; the backward pass reads x and dr once, writes dx and accumulates the 6 x Cin
// correlations dG[phi][delta] = sum dr[2m+phi] x[m+delta], from which a one-CTA finishing kernel forms the gradients of
// BOTH layers' kernels and biases (dV -> dWt = sum_j dV[j][k] (x) Wf[j], dWf[j] = sum_k Wt[k] dV[j][k] + dc_j bt, ...).
// Exact fp32 FMA arithmetic throughout; fixed-order reductions (deterministic).  HBM-bound: 61 MB forward, 119 MB backward
// at the benchmark shape instead of ~1.3 GB for the two layers run separately.
#include "common.cuh"

namespace vqb {

// gbuf (device, VQB_TAIL_GBUF floats): G[g][32] for g = phi*3 + (delta+1), then V00[32], V23[32], then K, c_0, c_2
constexpr int TG_V00 = 6 * 32, TG_V23 = 7 * 32, TG_K = 8 * 32, TG_C0 = TG_K + 1, TG_C2 = TG_K + 2;
static_assert(TG_C2 < VQB_TAIL_GBUF, "gbuf layout");
constexpr int TAIL_ROWS = 256;      // forward: input rows per CTA
constexpr int TAIL_PART = 7 * 32;   // backward: floats per CTA partial (6 correlations + sum of dr)

// which composed tap G[g] the product V[j][k] belongs to
__device__ __forceinline__ int tail_group(int j, int k) {
  // (0,0)->1 (0,1)->4 (0,2)->0 (0,3)->3 | (1,0)->5 (1,1)->1 (1,2)->4 (1,3)->0 | (2,0)->2 (2,1)->5 (2,2)->1 (2,3)->4
  const int tab[12] = {1, 4, 0, 3, 5, 1, 4, 0, 2, 5, 1, 4};
  return tab[j * 4 + k];
}

__global__ void __launch_bounds__(256) tail_prep_kernel(const float* __restrict__ wt, const float* __restrict__ bt,
                                                        const float* __restrict__ wf, const float* __restrict__ bf, int C,
                                                        int Cm, float* __restrict__ gbuf) {
  extern __shared__ __align__(16) float psm[];  // wt [4*Cm*C], wf [3*Cm], bt [Cm] staged once (coalesced), then V from shared memory
  float* wts = psm;
  float* wfs = wts + 4 * Cm * C;
  float* bts = wfs + 3 * Cm;
  __shared__ float V[12][32];
  __shared__ float cj[3];
  const int tid = threadIdx.x;
  for (int e = tid; e < 4 * Cm * C; e += blockDim.x) wts[e] = wt[e];
  for (int e = tid; e < 3 * Cm; e += blockDim.x) wfs[e] = wf[e];
  for (int e = tid; e < Cm; e += blockDim.x) bts[e] = bt ? bt[e] : 0.f;
  __syncthreads();
  for (int e = tid; e < 12 * 32; e += blockDim.x) {
    const int ci = e & 31, k = (e >> 5) & 3, j = e >> 7;
    float a = 0.f;
    if (ci < C)
      for (int c = 0; c < Cm; ++c) a = fmaf(wfs[j * Cm + c], wts[(k * Cm + c) * C + ci], a);
    V[j * 4 + k][ci] = a;
  }
  if (tid < 3) {
    float a = 0.f;
    for (int c = 0; c < Cm; ++c) a = fmaf(wfs[tid * Cm + c], bts[c], a);
    cj[tid] = a;
  }
  __syncthreads();
  for (int e = tid; e < 32; e += blockDim.x) {
    gbuf[0 * 32 + e] = V[0 * 4 + 2][e] + V[1 * 4 + 3][e];
    gbuf[1 * 32 + e] = (V[0 * 4 + 0][e] + V[1 * 4 + 1][e]) + V[2 * 4 + 2][e];
    gbuf[2 * 32 + e] = V[2 * 4 + 0][e];
    gbuf[3 * 32 + e] = V[0 * 4 + 3][e];
    gbuf[4 * 32 + e] = (V[0 * 4 + 1][e] + V[1 * 4 + 2][e]) + V[2 * 4 + 3][e];
    gbuf[5 * 32 + e] = V[1 * 4 + 0][e] + V[2 * 4 + 1][e];
    gbuf[TG_V00 + e] = V[0][e];
    gbuf[TG_V23 + e] = V[2 * 4 + 3][e];
  }
  if (tid == 0) {
    gbuf[TG_K] = ((bf ? bf[0] : 0.f) + cj[0]) + (cj[1] + cj[2]);
    gbuf[TG_C0] = cj[0];
    gbuf[TG_C2] = cj[2];
  }
}

// one thread per input position m: r[2m], r[2m+1] from rows m-1, m, m+1 of a shared-memory tile (row stride C+1: no conflicts)
__global__ void __launch_bounds__(TAIL_ROWS) tail_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gbuf,
                                                             float* __restrict__ r, int L, int C) {
  extern __shared__ __align__(16) float sm[];
  float* gs = sm;             // [8][32] composed taps + V00 + V23
  float* xs = sm + 8 * 32;    // [TAIL_ROWS + 2][C + 1]
  const int tid = threadIdx.x, b = blockIdx.y, m0 = blockIdx.x * TAIL_ROWS;
  const int c4n = C >> 2, ld = C + 1;
  gs[tid] = gbuf[tid];  // 8 * 32 == TAIL_ROWS
  const float* xb = x + (size_t)b * L * C;
  for (int e = tid; e < (TAIL_ROWS + 2) * c4n; e += TAIL_ROWS) {
    const int row = e / c4n, c4 = e - row * c4n;
    const int g = m0 - 1 + row;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g >= 0 && g < L) v = *reinterpret_cast<const float4*>(xb + (size_t)g * C + c4 * 4);
    float* d = xs + row * ld + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int m = m0 + tid;
  if (m >= L) return;
  const float K = gbuf[TG_K];
  float a0 = K, a1 = K;
  const float* xl = xs + tid * ld;
  const float* xc = xl + ld;
  const float* xr = xc + ld;
#pragma unroll 4
  for (int ci = 0; ci < C; ++ci) {
    const float l = xl[ci], c = xc[ci], rr = xr[ci];
    a0 = fmaf(l, gs[0 * 32 + ci], a0); a0 = fmaf(c, gs[1 * 32 + ci], a0); a0 = fmaf(rr, gs[2 * 32 + ci], a0);
    a1 = fmaf(l, gs[3 * 32 + ci], a1); a1 = fmaf(c, gs[4 * 32 + ci], a1); a1 = fmaf(rr, gs[5 * 32 + ci], a1);
  }
  if (m == 0) {  // y[-1] is padding, not the transposed convolution's formula value
    float s = gbuf[TG_C0];
    for (int ci = 0; ci < C; ++ci) s = fmaf(xc[ci], gs[TG_V00 + ci], s);
    a0 -= s;
  }
  if (m == L - 1) {  // y[2L] likewise
    float s = gbuf[TG_C2];
    for (int ci = 0; ci < C; ++ci) s = fmaf(xc[ci], gs[TG_V23 + ci], s);
    a1 -= s;
  }
  *reinterpret_cast<float2*>(r + ((size_t)b * L + m) * 2) = make_float2(a0, a1);
}

// one warp per run of rows, lane = input channel: dx row and the running correlations share the 6 dr coefficients
__global__ void __launch_bounds__(256) tail_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dr,
                                                       const float* __restrict__ gbuf, float* __restrict__ dx,
                                                       float* __restrict__ partial, long rows, int L, int C, long rpw) {
  __shared__ float red[8][TAIL_PART];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool act = lane < C;
  float G[6], v00, v23;
#pragma unroll
  for (int g = 0; g < 6; ++g) G[g] = gbuf[g * 32 + lane];
  v00 = gbuf[TG_V00 + lane];
  v23 = gbuf[TG_V23 + lane];
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float sdr = 0.f;
  const long r0 = ((long)blockIdx.x * 8 + warp) * rpw;
  const long r1 = r0 + rpw < rows ? r0 + rpw : rows;
  const int L2 = 2 * L;
  long b = r0 / L;
  int m = (int)(r0 - b * L) - 1;
#pragma unroll 4
  for (long row = r0; row < r1; ++row) {
    if (++m == L) { m = 0; ++b; }
    const int t = 2 * m - 2 + lane;  // lanes 0..5 fetch dr[2m-2 .. 2m+3]
    const float val = (lane < 6 && t >= 0 && t < L2) ? dr[b * L2 + t] : 0.f;
    const float xv = act ? x[row * C + lane] : 0.f;
    const float d0 = __shfl_sync(0xffffffffu, val, 0), d1 = __shfl_sync(0xffffffffu, val, 1);
    const float d2 = __shfl_sync(0xffffffffu, val, 2), d3 = __shfl_sync(0xffffffffu, val, 3);
    const float d4 = __shfl_sync(0xffffffffu, val, 4), d5 = __shfl_sync(0xffffffffu, val, 5);
    // coefficient of G[phi][delta] at this row is dr[2(m - delta) + phi]
    float o = d4 * G[0];
    o = fmaf(d2, G[1], o); o = fmaf(d0, G[2], o); o = fmaf(d5, G[3], o); o = fmaf(d3, G[4], o); o = fmaf(d1, G[5], o);
    if (m == 0) o = fmaf(-d2, v00, o);
    if (m == L - 1) o = fmaf(-d3, v23, o);
    if (act && dx) dx[row * C + lane] = o;
    acc[0] = fmaf(d4, xv, acc[0]); acc[1] = fmaf(d2, xv, acc[1]); acc[2] = fmaf(d0, xv, acc[2]);
    acc[3] = fmaf(d5, xv, acc[3]); acc[4] = fmaf(d3, xv, acc[4]); acc[5] = fmaf(d1, xv, acc[5]);
    sdr += (lane == 2 || lane == 3) ? val : 0.f;
  }
#pragma unroll
  for (int g = 0; g < 6; ++g) red[warp][g * 32 + lane] = acc[g];
  red[warp][6 * 32 + lane] = sdr;
  __syncthreads();
  if (tid < TAIL_PART) {
    float s = red[0][tid];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][tid];
    partial[(size_t)blockIdx.x * TAIL_PART + tid] = s;
  }
}

__global__ void __launch_bounds__(1024) tail_finish_kernel(const float* __restrict__ partial, int nparts,
                                                          const float* __restrict__ x, const float* __restrict__ dr,
                                                          const float* __restrict__ wt, const float* __restrict__ bt,
                                                          const float* __restrict__ wf, int B, int L, int C, int Cm,
                                                          float* __restrict__ dwt, float* __restrict__ dbt,
                                                          float* __restrict__ dwf, float* __restrict__ dbf) {
  __shared__ float dG[TAIL_PART];
  __shared__ float dGq[4][TAIL_PART];  // quarter sums of the CTA partials
  __shared__ float dV[12][32];
  __shared__ float b0[32], b1[32], sc[8];  // sc: sdr, e0, e1, dc_0..2
  const int tid = threadIdx.x;
  // weights staged once (coalesced): the gradient formulas below re-read them Cin * 4 times per output
  extern __shared__ __align__(16) float fsm[];
  float* wts = fsm;                 // [4*Cm*C]
  float* wfs = wts + 4 * Cm * C;    // [3*Cm]
  float* bts = wfs + 3 * Cm;        // [Cm]
  for (int e = tid; e < 4 * Cm * C; e += blockDim.x) wts[e] = wt[e];
  for (int e = tid; e < 3 * Cm; e += blockDim.x) wfs[e] = wf[e];
  for (int e = tid; e < Cm; e += blockDim.x) bts[e] = bt ? bt[e] : 0.f;
  {  // fixed-order sum over the CTA partials: 4 thread groups take a quarter of the parts each (8 independent chains per
     // thread, coalesced across threads), then the quarters are added in order: deterministic, 4x shorter than one group
    const int grp = tid >> 8, col = tid & 255;
    if (grp < 4 && col < TAIL_PART) {
      const int per = (nparts + 3) / 4, p0 = grp * per, p1 = min(nparts, p0 + per);
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int p = p0;
      for (; p + 8 <= p1; p += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] += partial[(size_t)(p + u) * TAIL_PART + col];
      }
      for (; p < p1; ++p) a[0] += partial[(size_t)p * TAIL_PART + col];
      dGq[grp][col] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    }
  }
  __syncthreads();
  if (tid < TAIL_PART) dG[tid] = (dGq[0][tid] + dGq[1][tid]) + (dGq[2][tid] + dGq[3][tid]);
  // boundary rows: the terms the zero padding of y removes at t = 0 and t = 2L-1
  if (tid < 32) {
    float s0 = 0.f, s1 = 0.f;
    if (tid < C)
      for (int b = 0; b < B; ++b) {
        s0 = fmaf(dr[(size_t)b * 2 * L], x[(size_t)b * L * C + tid], s0);
        s1 = fmaf(dr[(size_t)b * 2 * L + 2 * L - 1], x[((size_t)b * L + L - 1) * C + tid], s1);
      }
    b0[tid] = s0; b1[tid] = s1;
  }
  if (tid == 32 || tid == 33) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dr[(size_t)b * 2 * L + (tid == 32 ? 0 : 2 * L - 1)];
    sc[tid - 31] = s;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int l = 0; l < 32; ++l) s += dG[6 * 32 + l];
    sc[0] = s;
    sc[3] = s - sc[1]; sc[4] = s; sc[5] = s - sc[2];
  }
  for (int e = tid; e < 12 * 32; e += blockDim.x) {
    const int ci = e & 31, k = (e >> 5) & 3, j = e >> 7;
    float v = dG[tail_group(j, k) * 32 + ci];
    if (j == 0 && k == 0) v -= b0[ci];
    if (j == 2 && k == 3) v -= b1[ci];
    dV[j * 4 + k][ci] = v;
  }
  __syncthreads();
  for (int e = tid; e < 4 * Cm * C; e += blockDim.x) {  // dWt[k][c][ci]
    const int ci = e % C, c = (e / C) % Cm, k = e / (C * Cm);
    float s = dV[0 * 4 + k][ci] * wfs[0 * Cm + c];
    s = fmaf(dV[1 * 4 + k][ci], wfs[1 * Cm + c], s);
    s = fmaf(dV[2 * 4 + k][ci], wfs[2 * Cm + c], s);
    dwt[e] = s;
  }
  for (int e = tid; e < 3 * Cm; e += blockDim.x) {  // dWf[j][c]
    const int c = e % Cm, j = e / Cm;
    float s = sc[3 + j] * bts[c];
    for (int k = 0; k < 4; ++k)
      for (int ci = 0; ci < C; ++ci) s = fmaf(dV[j * 4 + k][ci], wts[(k * Cm + c) * C + ci], s);
    dwf[e] = s;
  }
  if (dbt)
    for (int c = tid; c < Cm; c += blockDim.x)
      dbt[c] = fmaf(sc[5], wfs[2 * Cm + c], fmaf(sc[4], wfs[1 * Cm + c], sc[3] * wfs[0 * Cm + c]));
  if (dbf && tid == 0) dbf[0] = sc[0];
}

static int tail_check(const vqb_tail_desc* d) {
  VQB_REQUIRE(d != nullptr, "null descriptor");
  VQB_REQUIRE(d->B >= 0 && d->L >= 0 && d->C_mid >= 1, "decoder tail: bad shape B=%d L=%d C_mid=%d", d->B, d->L, d->C_mid);
  VQB_REQUIRE(d->C_in >= 4 && d->C_in <= 32 && (d->C_in & 3) == 0,
              "decoder tail: C_in must be a multiple of 4 in [4, 32] (got %d); run the two layers separately", d->C_in);
  return VQB_OK;
}

static int tail_bwd_grid(const vqb_tail_desc* d, long* rpw) {
  const long rows = (long)d->B * d->L;
  long warps = cdiv(rows, 64);  // >= 64 rows per warp amortise its epilogue
  if (warps > 592 * 8) warps = 592 * 8;
  if (warps < 1) warps = 1;
  const int grid = cdiv(warps, 8);
  *rpw = cdiv(rows, (long)grid * 8);
  if (*rpw < 1) *rpw = 1;
  return grid;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_dec_tail_supports(const vqb_tail_desc* d) {
  return d && d->C_in >= 4 && d->C_in <= 32 && (d->C_in & 3) == 0 && d->C_mid >= 1 ? 1 : 0;
}

int vqb_dec_tail_fwd(const vqb_tail_desc* d, const float* x, const float* wt, const float* bt, const float* wf,
                     const float* bf, float* gbuf, float* recon, void* stream) {
  VQB_ARCH();
  if (int rc = tail_check(d)) return rc;
  VQB_REQUIRE(wt && wf && gbuf, "decoder tail: null weight / gbuf pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t psm = ((size_t)4 * d->C_mid * d->C_in + 4 * d->C_mid) * sizeof(float);
  VQB_REQUIRE(psm <= 48 * 1024, "decoder tail: C_mid = %d too wide for the weight staging buffer", d->C_mid);
  tail_prep_kernel<<<1, 256, psm, st>>>(wt, bt, wf, bf, d->C_in, d->C_mid, gbuf);
  VQB_LAUNCH_CHECK();
  if (d->B == 0 || d->L == 0) return VQB_OK;
  VQB_REQUIRE(x && recon, "decoder tail: null activation pointer");
  const size_t smem = (8 * 32 + (size_t)(TAIL_ROWS + 2) * (d->C_in + 1)) * sizeof(float);
  dim3 grid(cdiv(d->L, TAIL_ROWS), d->B);
  tail_fwd_kernel<<<grid, TAIL_ROWS, smem, st>>>(x, gbuf, recon, d->L, d->C_in);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

size_t vqb_dec_tail_bwd_workspace_bytes(const vqb_tail_desc* d) {
  long rpw;
  return (size_t)tail_bwd_grid(d, &rpw) * TAIL_PART * sizeof(float) + 64;
}

int vqb_dec_tail_bwd(const vqb_tail_desc* d, const float* x, const float* drecon, const float* wt, const float* bt,
                     const float* wf, const float* gbuf, float* dx, float* dwt, float* dbt, float* dwf, float* dbf,
                     void* workspace, size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  if (int rc = tail_check(d)) return rc;
  VQB_REQUIRE(wt && wf && gbuf && dwt && dwf, "decoder tail backward: null pointer");
  const size_t need = vqb_dec_tail_bwd_workspace_bytes(d);
  if (!workspace || workspace_bytes < need)
    return set_err(VQB_ERR_WORKSPACE, "decoder tail workspace: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  long rpw;
  const int grid = tail_bwd_grid(d, &rpw);
  const long rows = (long)d->B * d->L;
  float* partial = (float*)workspace;
  if (rows > 0) VQB_REQUIRE(x && drecon, "decoder tail backward: null activation pointer");
  tail_bwd_kernel<<<grid, 256, 0, st>>>(x, drecon, gbuf, dx, partial, rows, d->L > 0 ? d->L : 1, d->C_in, rpw);
  VQB_LAUNCH_CHECK();
  const size_t fsm = ((size_t)4 * d->C_mid * d->C_in + 4 * d->C_mid) * sizeof(float);
  VQB_REQUIRE(fsm <= 48 * 1024, "decoder tail: C_mid = %d too wide for the weight staging buffer", d->C_mid);
  tail_finish_kernel<<<1, 1024, fsm, st>>>(partial, grid, x, drecon, wt, bt, wf, d->L > 0 ? d->B : 0, d->L, d->C_in, d->C_mid,
                                        dwt, dbt, dwf, dbf);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

}  // extern "C"
