// wgrad_tc.cu — tensor-core weight gradient of the k=3, stride-1, 32->32 convolutions (the 208 residual-block
// convolutions of SMALL_VQ_VAE):   dW[j][ci][co] = sum_{b,t} act(x)[b, t + (j-1)*dil, ci] * dy[b, t, co],
//                                  dbias[co]     = sum_{b,t} dy[b, t, co].
// The reduction over time is the MMA K dimension: per tile of TK time rows both operands are staged once in the
// "plane layout" of tc.cuh and consumed as MN-major operands (rows = time = K):
//   A (M side)  = act(x) tile shifted by (j-1)*dil rows for tap j.  Its S bf16 pieces are consecutive plane groups, i.e.
//                 consecutive 32-row blocks of ONE M = 128 operand (row sx*32 + ci), followed by a constant plane whose
//                 first channel is 1.0: accumulator row 32*S is then the bias gradient;
//   B (N side)  = dy tile, its S pieces stacked the same way along N = 32*S (column sy*32 + co).
// One MMA per tap and K step therefore forms ALL S*S piece products; the epilogue adds the S*S accumulator blocks.
// Accumulators (3 taps x fp32 [128 x 32 S]) stay in TMEM for ALL tiles a CTA processes (persistent CTAs); one partial
// result per CTA goes to the workspace and is reduced in a fixed order (deterministic).
#include "common.cuh"
#include "tc.cuh"
#include <stdlib.h>

namespace vqb {

using namespace tc;

struct WgTcParams {
  const float* ga;  // gather side  [B, L, 32]  (x)
  const float* ot;  // other side   [B, L, 32]  (dy)
  float* partial;   // [CTAs of this problem][3*32*32 + 32]
  int B, L, dil, relu_ga, tiles_per_b, total_tiles;
  // Further problems of the same [B, L] (the other convolution of a residual block, the blocks of a whole DilatedResnet1D:
  // vqb_resblock_wgrad / vqb_resblock_wgrad_batch): the grid is cut into nprob equal runs of CTAs, run q works on problem q,
  // so that one launch (one prologue, one drain) yields all the weight gradients.  Problem 0 is the fields above.
  static constexpr int MAXP = 8;
  const float* gaq[MAXP];
  const float* otq[MAXP];
  float* partialq[MAXP];
  int dilq[MAXP], reluq[MAXP];
  int nprob;
  // The fixed-order sum over a problem's per-CTA partials happens inside the launch: the CTA that finishes a problem last
  // (counter in the workspace, zeroed by a memset node before the launch) adds the partials in the order of reduce_chunks_kernel
  // and writes dw / dbias — no separate reduction launches (they were 1.2 % of a training step).
  float* dwq[MAXP];
  float* dbq[MAXP];
  unsigned* counters;
  long long* trace;  // TRACE build (tools/trace_wgrad.py): clock64 stamps of CTA 0's tiles 6..9, [4][8]
};

#define WG_TR(ev)                                                                                                        \
  do {                                                                                                                   \
    if (TRACE && pp.trace && blockIdx.x == 0 && it >= 6 && it < 10 && (threadIdx.x & 31) == 0) pp.trace[(it - 6) * 8 + (ev)] = clock64(); \
  } while (0)

// S = number of bf16 pieces each operand is split into (1: bf16, 2: bf16x2, 3: bf16x3 = fp32-grade products)
template <int S_, int NT_ = 256>
struct WgCfg {
  static constexpr int S = S_;
  static constexpr int NT = NT_;                  // converter threads; one more warp issues the MMAs
  static constexpr int KMMA = 16;                 // K per MMA
  static constexpr int TK = 128;                  // time rows per tile
  static constexpr int DMAX = 32;
  static constexpr int NS = 32 * S;               // MMA N; live accumulator rows
  static constexpr int PLANE_D = TK * 16 + 32;                 // dy planes (8 bf16 channels x TK rows)
  static constexpr int PLANE_X = (TK + 2 * DMAX) * 16 + 32;    // act(x) planes (with the dilation halo)
  static constexpr int TILE_X = (4 * S + 1) * PLANE_X;         // S x 4 data planes + the constant ones plane
  static constexpr int TILE_D = 4 * S * PLANE_D;
  static constexpr int BUF = TILE_X + TILE_D;
  static constexpr int SPAN = 16 * PLANE_X;       // bytes an M = 128 A descriptor may touch from its start
  static constexpr int SMEM = (2 * BUF > BUF + SPAN ? 2 * BUF : BUF + SPAN) + 128;
  static constexpr int TCOLS = 3 * NS <= 128 ? 128 : 3 * NS <= 256 ? 256 : 512;  // TMEM columns: 3 taps x NS
  static constexpr int PART = 3 * 32 * 32 + 32;
  static constexpr int RPAD = 33;                 // row stride of the epilogue's shared-memory transpose (bank-conflict free)
  static constexpr int RED = 3 * 32 * RPAD + 32;  // floats per x piece in that buffer
  static constexpr int NA = TK * 4 / NT;                          // 8-channel units of the dy tile per converter thread
  static constexpr int NB = ((TK + 2 * DMAX) * 4 + NT - 1) / NT;  // ... of the act(x) tile
  static constexpr int NSETS = NT_ > 256 ? 2 : 3;  // register sets: tiles i+1 (and i+2) are in flight while tile i is converted
};

// one tile's worth of global data held by a converter thread: its 8 channels of one dy row and of up to NB act(x) rows
template <int NA, int NB>
struct WgRegs {
  float4 a[NA][2];
  float4 b[NB][2];
};

// Pipeline: the converter warps (16 by default) keep NSETS-1 tiles of global loads in flight in registers, split the tile whose data
// has arrived into bf16 pieces in one of two shared-memory operand buffers and signal full[buf]; the issuing warp waits
// for full[buf], issues the tile's MMAs and commits them to empty[buf], which the converters wait on before they
// overwrite that buffer two tiles later.  Nobody waits for a global load it issued less than two tiles ago.
template <int S, int NT_, bool TRACE = false>
__global__ void __launch_bounds__(NT_ + 32, 1) wgrad_tc_kernel(const WgTcParams pp) {
  using Cfg = WgCfg<S, NT_>;
  // CTA-local view: which problem, which slice of its tiles
  const int nblk = (int)gridDim.x / pp.nprob;  // the grid is nprob * nblk CTAs
  const int prob = (int)blockIdx.x / nblk, bidx = (int)blockIdx.x - prob * nblk;
  struct { const float* ga; const float* ot; float* partial; int B, L, dil, relu_ga, tiles_per_b, total_tiles; } p;
  p.B = pp.B; p.L = pp.L; p.tiles_per_b = pp.tiles_per_b; p.total_tiles = pp.total_tiles;
  if (prob == 0) { p.ga = pp.ga; p.ot = pp.ot; p.partial = pp.partial; p.dil = pp.dil; p.relu_ga = pp.relu_ga; }
  else { p.ga = pp.gaq[prob]; p.ot = pp.otq[prob]; p.partial = pp.partialq[prob]; p.dil = pp.dilq[prob]; p.relu_ga = pp.reluq[prob]; }
  constexpr int NT = Cfg::NT, NA = Cfg::NA, NB = Cfg::NB, NSETS = Cfg::NSETS;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[2], empty[2], done;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(&tslot, Cfg::TCOLS);
  if (tid == 32) {
    mbar_init(&full[0], NT / 32); mbar_init(&full[1], NT / 32);
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); mbar_init(&done, 1);
    fence_mbar_init();
  }
  // constant plane of both act(x) tiles: channel 32*S = 1.0 in every row (bias row of the accumulator), the other 7 are 0
  for (int e = tid; e < 2 * (Cfg::PLANE_X / 16); e += NT + 32) {
    const int buf = e / (Cfg::PLANE_X / 16), r = e - buf * (Cfg::PLANE_X / 16);
    *reinterpret_cast<uint4*>(smem + buf * Cfg::BUF + 4 * S * Cfg::PLANE_X + r * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  pdl_launch_dependents();  // prologue done (common.cuh): the next kernel may become resident ...
  pdl_wait();               // ... and this one must not touch the previous kernel's outputs before it has completed

  const long first = (long)bidx * p.total_tiles / nblk;
  const long last = (long)(bidx + 1) * p.total_tiles / nblk;
  const int ntiles = (int)(last - first);
  const int rowsB = Cfg::TK + 2 * p.dil;

  if (warp == NT / 32) {
    // ---------------------------------------------------------------------------------- MMA issuer (one thread)
    if (elect_one()) {
      const uint32_t idesc = instr_desc(FMT_BF16, 128, Cfg::NS, true, true);
      const uint64_t dil = (uint64_t)p.dil;  // a row is 16 bytes = one unit of the descriptor's start-address field
      for (int it = 0; it < ntiles; ++it) {
        const int buf = it & 1;
        WG_TR(0);
        mbar_wait(&full[buf], (it >> 1) & 1);
        fence_after_sync();
        WG_TR(1);
        const uint64_t ad0 = smem_desc(smem_u32(smem + buf * Cfg::BUF), 128, Cfg::PLANE_X);
        const uint64_t bd0 = smem_desc(smem_u32(smem + buf * Cfg::BUF + Cfg::TILE_X), 128, Cfg::PLANE_D);
        uint32_t acc = it != 0;
#pragma unroll
        for (int ks = 0; ks < Cfg::TK / Cfg::KMMA; ++ks) {
          // tap j reads act(x) rows t + (j-1)*dil = x-tile rows ks*KMMA + j*dil
          const uint64_t ad = ad0 + (uint64_t)(ks * Cfg::KMMA), bd = bd0 + (uint64_t)(ks * Cfg::KMMA);
          mma<false>(tmem + 0 * Cfg::NS, ad, bd, idesc, acc);
          mma<false>(tmem + 1 * Cfg::NS, ad + dil, bd, idesc, acc);
          mma<false>(tmem + 2 * Cfg::NS, ad + 2 * dil, bd, idesc, acc);
          acc = 1;
        }
        commit(&empty[buf]);
        WG_TR(2);
      }
      commit(&done);  // everything issued so far
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------- loaders / converters
    const int o = tid & 3;  // 8-channel unit of this thread (the same for all its rows: idx = tid + k*NT, NT % 4 == 0)
    // tiles are loaded strictly in order: a running (batch item, tile in item) cursor replaces the 64-bit division per tile and
    // thread that tools/trace_wgrad.py showed on the converters' critical path (~600 clk per tile to issue six loads)
    int ld_b = (int)(first / p.tiles_per_b), ld_tx = (int)(first - (long)ld_b * p.tiles_per_b);
    auto load = [&](int it, WgRegs<NA, NB>& R) {
      const int b = ld_b;
      const int t0 = ld_tx * Cfg::TK;
      if (++ld_tx == p.tiles_per_b) { ld_tx = 0; ++ld_b; }
      const float* otb = p.ot + (size_t)b * p.L * 32 + o * 8;
      const float* gab = p.ga + (size_t)b * p.L * 32 + o * 8;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        const int g = t0 + ((tid + k * NT) >> 2);
        const bool ok = g < p.L;
        R.a[k][0] = ok ? *reinterpret_cast<const float4*>(otb + (size_t)g * 32) : z;
        R.a[k][1] = ok ? *reinterpret_cast<const float4*>(otb + (size_t)g * 32 + 4) : z;
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int r = (tid + k * NT) >> 2;
        const int g = t0 - p.dil + r;
        const bool ok = r < rowsB && g >= 0 && g < p.L;
        R.b[k][0] = ok ? *reinterpret_cast<const float4*>(gab + (long)g * 32) : z;
        R.b[k][1] = ok ? *reinterpret_cast<const float4*>(gab + (long)g * 32 + 4) : z;
      }
    };
    auto convert = [&](int it, WgRegs<NA, NB>& R) {
      const int buf = it & 1;
      uint8_t* Xt = smem + buf * Cfg::BUF;
      uint8_t* Dt = Xt + Cfg::TILE_X;
      if (warp == 0) WG_TR(3);
      // "the MMAs that read this buffer have completed": even a satisfied mbarrier wait returns only after ~150-400 clk (tools/mma_probe,
      // tools/trace_wgrad.py), and all converter warps reach it together — so the poll is issued first, the dy rows are split in
      // registers under its latency, and only the shared-memory stores wait for the answer
      const uint32_t epar = (uint32_t)(((it >> 1) - 1) & 1);
      const uint32_t free_now = it >= 2 ? mbar_try(&empty[buf], epar) : 1u;
      uint4 pa[NA][S];
#pragma unroll
      for (int k = 0; k < NA; ++k) split8<S>(R.a[k][0], R.a[k][1], pa[k]);
      if (!free_now) mbar_wait(&empty[buf], epar);
      if (warp == 0) WG_TR(4);
      uint4 pc[S];
#pragma unroll
      for (int k = 0; k < NA; ++k) {
#pragma unroll
        for (int s = 0; s < S; ++s)
          *reinterpret_cast<uint4*>(Dt + (s * 4 + o) * Cfg::PLANE_D + ((tid + k * NT) >> 2) * 16) = pa[k][s];
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int r = (tid + k * NT) >> 2;
        if (r < rowsB) {
          float4 u = R.b[k][0], w = R.b[k][1];
          if (p.relu_ga) {
            u.x = fmaxf(u.x, 0.f); u.y = fmaxf(u.y, 0.f); u.z = fmaxf(u.z, 0.f); u.w = fmaxf(u.w, 0.f);
            w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f);
          }
          split8<S>(u, w, pc);
#pragma unroll
          for (int s = 0; s < S; ++s) *reinterpret_cast<uint4*>(Xt + (s * 4 + o) * Cfg::PLANE_X + r * 16) = pc[s];
        }
      }
      if (warp == 0) WG_TR(5);
      fence_proxy_async();  // these generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[buf]);
      if (warp == 0) WG_TR(6);
    };
    WgRegs<NA, NB> R[NSETS];
#pragma unroll
    for (int u = 0; u < NSETS - 1; ++u)
      if (u < ntiles) load(u, R[u]);
#pragma unroll 1
    for (int base = 0; base < ntiles; base += NSETS) {
#pragma unroll
      for (int u = 0; u < NSETS; ++u) {
        const int it = base + u;
        if (it < ntiles) {
          if (it + NSETS - 1 < ntiles) load(it + NSETS - 1, R[(u + NSETS - 1) % NSETS]);
          if (warp == 0) WG_TR(7);
          convert(it, R[u]);
        }
      }
    }
  }

  float* out = p.partial + (size_t)bidx * Cfg::PART;
  if (ntiles == 0) {  // no tiles: zero partial
    for (int e = tid; e < Cfg::PART; e += NT + 32) out[e] = 0.f;
  } else {
    mbar_wait(&done, 0);
    fence_after_sync();
    float* red = reinterpret_cast<float*>(smem);  // [S][RED]: the operand tiles are dead now
    if (warp < S) {  // accumulator rows 32*warp .. +31: x piece `warp`, input channel ci = lane
      float* r = red + warp * Cfg::RED;
      const uint32_t ta = tmem + (((uint32_t)warp * 32u) << 16);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v[32], m[32];
        tmem_ld32(ta + j * Cfg::NS, v);
#pragma unroll
        for (int sy = 1; sy < S; ++sy) {  // add the column blocks of the other dy pieces
          tmem_ld32(ta + j * Cfg::NS + sy * 32, m);
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] += m[c];
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) r[(j * 32 + lane) * Cfg::RPAD + c] = v[c];
      }
    }
    if (warp == S) {  // accumulator row 32*S (lane 0 of this warp's quadrant; S < 4): the ones row = bias gradient
      float v[32], m[32];
      const uint32_t ta = tmem + (((uint32_t)S * 32u) << 16);
      tmem_ld32(ta + 1 * Cfg::NS, v);
#pragma unroll
      for (int sy = 1; sy < S; ++sy) {
        tmem_ld32(ta + 1 * Cfg::NS + sy * 32, m);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += m[c];
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 32; ++c) red[3 * 32 * Cfg::RPAD + c] = v[c];
      }
    }
    fence_before_sync();
    __syncthreads();
    for (int e = tid; e < Cfg::PART; e += NT + 32) {
      const int row = e >> 5, c = e & 31;  // row = j*32 + ci for the weights, 96 for the bias
      const int a = row * Cfg::RPAD + c;
      float t = red[a];
      if (row < 96) {
#pragma unroll
        for (int w = 1; w < S; ++w) t += red[w * Cfg::RED + a];
      }
      out[e] = t;
    }
  }
  if (pp.counters) {
    __shared__ int is_last;
    __threadfence();  // this CTA's partial is visible device-wide before it is counted
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(&pp.counters[prob], 1u) == (unsigned)(nblk - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      const float* base = p.partial;  // [nblk][PART]
      float* dw = pp.dwq[prob];
      float* db = pp.dbq[prob];
      // three elements x eight chains = 24 independent loads per step: a single CTA has to pull nblk x 12 KB out of L2, and
      // with one load in flight per thread that took longer than the separate reduction kernels it replaces
      constexpr int EB = 3;
      for (int e0 = tid; e0 < Cfg::PART; e0 += EB * (NT + 32)) {
        float sl[EB][8];
#pragma unroll
        for (int k = 0; k < EB; ++k)
#pragma unroll
          for (int l = 0; l < 8; ++l) sl[k][l] = 0.f;
        for (int c0 = 0; c0 < nblk; c0 += 8) {
          float v[EB][8];
#pragma unroll
          for (int k = 0; k < EB; ++k)
#pragma unroll
            for (int l = 0; l < 8; ++l) {
              const int e = e0 + k * (NT + 32);
              v[k][l] = (c0 + l < nblk && e < Cfg::PART) ? __ldcg(base + (size_t)(c0 + l) * Cfg::PART + e) : 0.f;
            }
#pragma unroll
          for (int k = 0; k < EB; ++k)
#pragma unroll
            for (int l = 0; l < 8; ++l) sl[k][l] += v[k][l];  // chain l: chunks l, l + 8, ... in order (+ 0.f past the end: exact)
        }
#pragma unroll
        for (int k = 0; k < EB; ++k) {
          const int e = e0 + k * (NT + 32);
          if (e < Cfg::PART) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < 8; ++l) t += sl[k][l];  // the summation tree of reduce_chunks_kernel
            if (e < 3 * 32 * 32) dw[e] = t;
            else if (db) db[e - 3 * 32 * 32] = t;
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

// in-kernel reduction (last CTA of a problem) only up to this many partials per problem: one CTA pulling 12 KB per partial out of
// L2 is quicker than the separate reduction launches for the 18 partials of a batched launch, not for the 74-148 of a small one
constexpr int WG_INKERNEL_MAX = 32;

static int wg_split(int precision) { return (precision == VQB_PREC_BF16X3 || precision == VQB_PREC_FP16X2) ? 3 : precision == VQB_PREC_BF16X2 ? 2 : 1; }

bool wgrad_tc_supported(const vqb_conv_desc* d) {
  // kind::tf32 with MN-major (time-as-K) operands returned zeros on B200, so VQB_PREC_TF32 has no weight-gradient kernel
  return d->k == 3 && d->stride == 1 && d->C_in == 32 && d->C_out == 32 && d->dilation >= 1 && d->dilation <= 32 &&
         (d->precision == VQB_PREC_BF16 || d->precision == VQB_PREC_BF16X2 || d->precision == VQB_PREC_BF16X3 ||
          d->precision == VQB_PREC_FP16X2);
}

static int wgrad_tc_grid(const vqb_conv_desc* d, int* tiles_per_b) {
  *tiles_per_b = cdiv(d->L, WgCfg<1>::TK);
  const long total = (long)d->B * *tiles_per_b;
  static int num_sms = 0;  // one persistent CTA per SM (the operand buffers take most of its shared memory)
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) num_sms = 148;
  }
  return (int)(total < num_sms ? (total > 0 ? total : 1) : num_sms);
}

size_t wgrad_tc_workspace_bytes(const vqb_conv_desc* d) {
  int tpb;
  return (size_t)wgrad_tc_grid(d, &tpb) * WgCfg<1>::PART * sizeof(float) + 64;
}

template <int S, int NT = 256>
static int launch_wg(const WgTcParams& p, int grid, cudaStream_t st) {
  using Cfg = WgCfg<S, NT>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<S, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  VQB_CUDA(launch_pdl(wgrad_tc_kernel<S, NT>, dim3(grid), dim3(Cfg::NT + 32), (size_t)Cfg::SMEM, st, p));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

static int launch_wg_any(int S, const WgTcParams& p, int grid, cudaStream_t st) {
  // converter threads: 512 (16 warps, 2 register sets: one tile of loads in flight ahead) measured 2.3 % faster per training
  // step than 256 (8 warps, 3 sets) once the launches were batched; VQB_WGRAD_NT=256 selects the other variant
  const char* e = getenv("VQB_WGRAD_NT");
  const int nt = e ? atoi(e) : 512;
  if (getenv("VQB_WG_TRACE") && S == 3 && nt == 512) {  // profiling aid (tools/trace_wgrad.py): device buffer of 32 int64
    WgTcParams q = p;
    q.trace = reinterpret_cast<long long*>(strtoull(getenv("VQB_WG_TRACE"), nullptr, 0));
    static bool set = false;
    if (!set) { VQB_CUDA((cudaFuncSetAttribute(wgrad_tc_kernel<3, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<3, 512>::SMEM))); set = true; }
    VQB_CUDA((launch_pdl(wgrad_tc_kernel<3, 512, true>, dim3(grid), dim3(512 + 32), (size_t)WgCfg<3, 512>::SMEM, st, q)));
    VQB_LAUNCH_CHECK();
    return VQB_OK;
  }
  return nt == 512 ? (S == 3 ? launch_wg<3, 512>(p, grid, st) : S == 2 ? launch_wg<2, 512>(p, grid, st) : launch_wg<1, 512>(p, grid, st))
                   : (S == 3 ? launch_wg<3>(p, grid, st) : S == 2 ? launch_wg<2>(p, grid, st) : launch_wg<1>(p, grid, st));
}

// The weight gradients of n residual blocks of the same [B, L, 32] (resnet.py:13-17; n = 1: one block, n = 4: a whole
// DilatedResnet1D) in ONE launch: problem 2i = conv1 of block i (act(x), dh, its dilation), problem 2i + 1 = conv2
// (act(h), dy, dilation 1).  d describes the shape (k 3, 32 -> 32); its dilation field is ignored.
size_t resblock_wgrad_tc_workspace_bytes(const vqb_conv_desc* d, int n) { return (size_t)2 * n * wgrad_tc_workspace_bytes(d); }

int resblock_wgrad_tc(const vqb_conv_desc* d, int n, const int* dilations, const float* const* x, const float* const* h,
                      const float* const* dy, const float* const* dh, float* const* dw1, float* const* db1,
                      float* const* dw2, float* const* db2, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n < 1 || 2 * n > WgTcParams::MAXP) return set_err(VQB_ERR_INVALID, "residual-block wgrad: 1..%d blocks per launch (got %d)", WgTcParams::MAXP / 2, n);
  const size_t need = resblock_wgrad_tc_workspace_bytes(d, n);
  if (!ws || ws_bytes < need) return set_err(VQB_ERR_WORKSPACE, "residual-block wgrad workspace: need %zu bytes, got %zu", need, ws_bytes);
  constexpr int PART = WgCfg<1>::PART;
  WgTcParams p{};
  const int np = 2 * n;
  int g = wgrad_tc_grid(d, &p.tiles_per_b);  // CTAs a single problem would get (<= number of SMs)
  p.B = d->B; p.L = d->L; p.total_tiles = d->B * p.tiles_per_b; p.nprob = np;
  const int full = g;
  g = full / np > 0 ? full / np : 1;         // ... per problem here
  if (g > p.total_tiles) g = p.total_tiles;
  for (int q = 0; q < np; ++q) {
    const int i = q >> 1;
    const bool c2 = q & 1;
    p.gaq[q] = c2 ? h[i] : x[i]; p.otq[q] = c2 ? dy[i] : dh[i]; p.dilq[q] = c2 ? 1 : dilations[i]; p.reluq[q] = 1;
    p.partialq[q] = (float*)ws + (size_t)q * g * PART;
  }
  p.ga = p.gaq[0]; p.ot = p.otq[0]; p.dil = p.dilq[0]; p.relu_ga = 1; p.partial = p.partialq[0];
  for (int q = 0; q < np; ++q) {
    p.dwq[q] = (q & 1) ? dw2[q >> 1] : dw1[q >> 1];
    p.dbq[q] = (q & 1) ? db2[q >> 1] : db1[q >> 1];
  }
  if (g <= WG_INKERNEL_MAX) {  // few partials per problem (a batched launch): the last CTA of each problem adds them up
    p.counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + need - 64);  // the spare tail of the workspace
    VQB_CUDA(cudaMemsetAsync(p.counters, 0, WgTcParams::MAXP * sizeof(unsigned), st));
    return launch_wg_any(wg_split(d->precision), p, np * g, st);
  }
  int rc = launch_wg_any(wg_split(d->precision), p, np * g, st);  // many partials: the parallel reduction kernels are faster than one CTA
  if (rc) return rc;
  for (int q = 0; q < np; ++q) {
    reduce_chunks_strided(p.partialq[q], g, PART, 0, 3 * 32 * 32, p.dwq[q], st);
    VQB_LAUNCH_CHECK();
    if (p.dbq[q]) {
      reduce_chunks_strided(p.partialq[q], g, PART, 3 * 32 * 32, 32, p.dbq[q], st);
      VQB_LAUNCH_CHECK();
    }
  }
  return VQB_OK;
}

int conv1d_wgrad_tc(const vqb_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  const size_t need = wgrad_tc_workspace_bytes(d);
  if (!ws || ws_bytes < need) return set_err(VQB_ERR_WORKSPACE, "tensor-core wgrad workspace: need %zu bytes, got %zu", need, ws_bytes);
  WgTcParams p{};
  p.ga = x; p.ot = dy; p.partial = (float*)ws;
  p.B = d->B; p.L = d->L; p.dil = d->dilation; p.relu_ga = d->relu_in;
  const int grid = wgrad_tc_grid(d, &p.tiles_per_b);
  p.total_tiles = d->B * p.tiles_per_b;
  const int S = wg_split(d->precision);
  p.nprob = 1;
  p.dwq[0] = dw; p.dbq[0] = dbias;
  if (grid <= WG_INKERNEL_MAX) {
    p.counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + need - 64);
    VQB_CUDA(cudaMemsetAsync(p.counters, 0, WgTcParams::MAXP * sizeof(unsigned), st));
    return launch_wg_any(S, p, grid, st);
  }
  int rc = launch_wg_any(S, p, grid, st);
  if (rc) return rc;
  constexpr int PART = WgCfg<1>::PART;
  reduce_chunks_strided(p.partial, grid, PART, 0, 3 * 32 * 32, dw, st);
  VQB_LAUNCH_CHECK();
  if (dbias) {
    reduce_chunks_strided(p.partial, grid, PART, 3 * 32 * 32, 32, dbias, st);
    VQB_LAUNCH_CHECK();
  }
  return VQB_OK;
}

}  // namespace vqb
