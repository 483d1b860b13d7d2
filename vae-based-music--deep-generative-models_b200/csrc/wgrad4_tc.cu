// wgrad4_tc.cu — tensor-core weight gradient of the k = 4, stride-2, 32 <-> 32 convolutions:
//     dW[n][cg][co] = sum_{b,m} ga[b, 2m + n - 1, cg] * ot[b, m, co],   n = 0..3   (rows outside the tensor count as 0)
//   Conv1D(32, 4, strides=2)          (encdec.py:33):    ga = x  [B, L, 32],   ot = dy [B, ceil(L/2), 32] -> dW [k, Cin, Cout],
//                                                        dbias = column sums of ot
//   Conv1DTranspose(32, 4, strides=2) (encdec.py:67-68): ga = dy [B, 2L, 32],  ot = x  [B, L, 32]         -> dW [k, Cout, Cin],
//                                                        dbias = column sums of ga
// Same scheme as wgrad_tc.cu (time is the MMA K dimension, both operands MN-major in the plane layout, S bf16 pieces of
// ga stacked along M and of ot along N, one MMA per tap and K step forming all piece products, accumulators resident in
// TMEM over all tiles of a persistent CTA).  The stride is removed at staging time: ga rows are split by parity into an
// even and an odd operand tile, after which tap n reads tile (n + 1) % 2 shifted by (n + 1) / 2 rows:
//     n = 0: odd[m-1]   n = 1: even[m]   n = 2: odd[m]   n = 3: even[m+1].
// Bias gradients ride along: a constant-ones channel behind the ga pieces makes accumulator row 32 S the column sums of ot;
// a constant-ones channel behind the ot pieces makes accumulator column 32 S of taps 1 and 2 the column sums of the even
// and the odd ga rows.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace vqb {

using namespace tc;

struct Wg4Params {
  const float* ga;  // [B, Lg, 32]
  const float* ot;  // [B, Lo, 32]
  float* partial;   // [gridDim.x][PART]
  int B, Lg, Lo, tiles_per_b, total_tiles;
};

template <int S_>
struct Wg4Cfg {
  static constexpr int S = S_;
  static constexpr int NT = 256;                   // converter threads; one more warp issues the MMAs
  static constexpr int KMMA = 16;
  static constexpr int TK = 128;                   // ot rows per tile
  static constexpr int GROWS = TK + 2;             // rows of each parity tile of ga (one guard row each side)
  static constexpr int NS = 32 * S;
  static constexpr int NCOL = NS + 8;              // MMA N: ot pieces + the ones channel (N must be a multiple of 8)
  static constexpr int PLANE_G = GROWS * 16 + 32;
  static constexpr int PLANE_O = TK * 16 + 32;
  static constexpr int TILE_G = (4 * S + 1) * PLANE_G;   // S x 4 data planes + the ones plane (one parity)
  static constexpr int TILE_O = (4 * S + 1) * PLANE_O;
  static constexpr int BUF = 2 * TILE_G + TILE_O;
  static constexpr int SPAN = 16 * PLANE_G;        // bytes an M = 128 A descriptor may touch from its start
  static constexpr int SMEM = 2 * BUF + SPAN + 128;
  static constexpr int TCOLS = 512;                // 4 taps x NCOL columns
  static constexpr int PART = 4 * 32 * 32 + 64;    // dW, column sums of ot, column sums of ga
  static constexpr int RPAD = 33;
  static constexpr int RED = 4 * 32 * RPAD + 64;   // floats per ga piece in the epilogue's transpose buffer
  static constexpr int NG = (2 * GROWS * 4 + NT - 1) / NT;  // 8-channel units of the ga rows per converter thread
  static constexpr int NO = TK * 4 / NT;                    // ... of the ot rows
  static constexpr int NSETS = 2;                  // register sets: tile i+1 is in flight while tile i is converted
  static_assert(4 * NCOL <= TCOLS, "TMEM columns");
};

template <int NG, int NO>
struct Wg4Regs {
  float4 g[NG][2];
  float4 o[NO][2];
};

template <int S>
__global__ void __launch_bounds__(Wg4Cfg<S>::NT + 32, 1) wgrad4_tc_kernel(const Wg4Params p) {
  using Cfg = Wg4Cfg<S>;
  constexpr int NT = Cfg::NT, NG = Cfg::NG, NO = Cfg::NO, NSETS = Cfg::NSETS;
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
  // constant ones planes (first channel 1.0, the other 7 zero) behind the data planes of every operand tile
  for (int e = tid; e < 2 * (2 * (Cfg::PLANE_G / 16) + Cfg::PLANE_O / 16); e += NT + 32) {
    const int per = 2 * (Cfg::PLANE_G / 16) + Cfg::PLANE_O / 16;
    const int buf = e / per;
    int r = e - buf * per;
    uint8_t* base = smem + buf * Cfg::BUF;
    uint8_t* dst;
    if (r < Cfg::PLANE_G / 16) dst = base + 4 * S * Cfg::PLANE_G + r * 16;
    else if ((r -= Cfg::PLANE_G / 16) < Cfg::PLANE_G / 16) dst = base + Cfg::TILE_G + 4 * S * Cfg::PLANE_G + r * 16;
    else dst = base + 2 * Cfg::TILE_G + 4 * S * Cfg::PLANE_O + (r - Cfg::PLANE_G / 16) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(0x00003F80u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  pdl_launch_dependents();
  pdl_wait();

  const long first = (long)blockIdx.x * p.total_tiles / gridDim.x;
  const long last = (long)(blockIdx.x + 1) * p.total_tiles / gridDim.x;
  const int ntiles = (int)(last - first);

  if (warp == NT / 32) {
    // ---------------------------------------------------------------------------------- MMA issuer (one thread)
    if (elect_one()) {
      const uint32_t idesc = instr_desc(FMT_BF16, 128, Cfg::NCOL, true, true);
      for (int it = 0; it < ntiles; ++it) {
        const int buf = it & 1;
        mbar_wait(&full[buf], (it >> 1) & 1);
        fence_after_sync();
        const uint64_t ae = smem_desc(smem_u32(smem + buf * Cfg::BUF), 128, Cfg::PLANE_G);              // even ga rows
        const uint64_t ao = smem_desc(smem_u32(smem + buf * Cfg::BUF + Cfg::TILE_G), 128, Cfg::PLANE_G);  // odd ga rows
        const uint64_t bd0 = smem_desc(smem_u32(smem + buf * Cfg::BUF + 2 * Cfg::TILE_G), 128, Cfg::PLANE_O);
        uint32_t acc = it != 0;
#pragma unroll
        for (int ks = 0; ks < Cfg::TK / Cfg::KMMA; ++ks) {
          // tile row r of a parity tile holds u = m0 - 1 + r: ot row m0 + i pairs with rows i (u = m-1), i+1 (u = m), i+2 (u = m+1)
          const uint64_t k = (uint64_t)(ks * Cfg::KMMA), bd = bd0 + k;
          mma<false>(tmem + 0 * Cfg::NCOL, ao + k + 0, bd, idesc, acc);  // n = 0: odd[m-1]
          mma<false>(tmem + 1 * Cfg::NCOL, ae + k + 1, bd, idesc, acc);  // n = 1: even[m]
          mma<false>(tmem + 2 * Cfg::NCOL, ao + k + 1, bd, idesc, acc);  // n = 2: odd[m]
          mma<false>(tmem + 3 * Cfg::NCOL, ae + k + 2, bd, idesc, acc);  // n = 3: even[m+1]
          acc = 1;
        }
        commit(&empty[buf]);
      }
      commit(&done);
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------- loaders / converters
    const int o = tid & 3;  // 8-channel unit of this thread
    // tiles are loaded strictly in order: a running (batch item, tile in item) cursor instead of a 64-bit division per tile and thread
    int ld_b = (int)(first / p.tiles_per_b), ld_tx = (int)(first - (long)ld_b * p.tiles_per_b);
    auto load = [&](int it, Wg4Regs<NG, NO>& R) {
      const int b = ld_b;
      const int m0 = ld_tx * Cfg::TK;
      if (++ld_tx == p.tiles_per_b) { ld_tx = 0; ++ld_b; }
      const float* otb = p.ot + (size_t)b * p.Lo * 32 + o * 8;
      const float* gab = p.ga + (size_t)b * p.Lg * 32 + o * 8;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < NO; ++k) {
        const int m = m0 + ((tid + k * NT) >> 2);
        const bool ok = m < p.Lo;
        R.o[k][0] = ok ? *reinterpret_cast<const float4*>(otb + (size_t)m * 32) : z;
        R.o[k][1] = ok ? *reinterpret_cast<const float4*>(otb + (size_t)m * 32 + 4) : z;
      }
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const int rr = (tid + k * NT) >> 2;      // staged ga row: g = 2 (m0 - 1) + rr, parity rr & 1, tile row rr >> 1
        const int g = 2 * (m0 - 1) + rr;
        const bool ok = rr < 2 * Cfg::GROWS && g >= 0 && g < p.Lg;
        R.g[k][0] = ok ? *reinterpret_cast<const float4*>(gab + (long)g * 32) : z;
        R.g[k][1] = ok ? *reinterpret_cast<const float4*>(gab + (long)g * 32 + 4) : z;
      }
    };
    auto convert = [&](int it, Wg4Regs<NG, NO>& R) {
      const int buf = it & 1;
      uint8_t* Gt = smem + buf * Cfg::BUF;
      uint8_t* Ot = Gt + 2 * Cfg::TILE_G;
      if (it >= 2) mbar_wait(&empty[buf], ((it >> 1) - 1) & 1);  // the MMAs that read this buffer have completed
      uint4 pc[S];
#pragma unroll
      for (int k = 0; k < NO; ++k) {
        split8<S>(R.o[k][0], R.o[k][1], pc);
#pragma unroll
        for (int s = 0; s < S; ++s)
          *reinterpret_cast<uint4*>(Ot + (s * 4 + o) * Cfg::PLANE_O + ((tid + k * NT) >> 2) * 16) = pc[s];
      }
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const int rr = (tid + k * NT) >> 2;
        if (rr < 2 * Cfg::GROWS) {
          split8<S>(R.g[k][0], R.g[k][1], pc);
#pragma unroll
          for (int s = 0; s < S; ++s)
            *reinterpret_cast<uint4*>(Gt + (rr & 1) * Cfg::TILE_G + (s * 4 + o) * Cfg::PLANE_G + (rr >> 1) * 16) = pc[s];
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[buf]);
    };
    Wg4Regs<NG, NO> R[NSETS];
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
          convert(it, R[u]);
        }
      }
    }
  }

  float* out = p.partial + (size_t)blockIdx.x * Cfg::PART;
  if (ntiles == 0) {
    for (int e = tid; e < Cfg::PART; e += NT + 32) out[e] = 0.f;
  } else {
    mbar_wait(&done, 0);
    fence_after_sync();
    float* red = reinterpret_cast<float*>(smem);  // [S][RED]: the operand tiles are dead now
    if (warp < S) {  // accumulator rows 32*warp .. +31: ga piece `warp`, channel cg = lane
      float* r = red + warp * Cfg::RED;
      const uint32_t ta = tmem + (((uint32_t)warp * 32u) << 16);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        float v[32], m[32];
        tmem_ld32(ta + n * Cfg::NCOL, v);
#pragma unroll
        for (int sy = 1; sy < S; ++sy) {  // add the column blocks of the other ot pieces
          tmem_ld32(ta + n * Cfg::NCOL + sy * 32, m);
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] += m[c];
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) r[(n * 32 + lane) * Cfg::RPAD + c] = v[c];
      }
      // column 32 S of taps 1 and 2: sums over the even / odd ga rows of this piece's channel cg = lane
      float e8[16], o8[16];
      tmem_ld16(ta + 1 * Cfg::NCOL + Cfg::NS - 8, e8);  // 16 columns ending with the ones column block: [NS-8, NS+8)
      tmem_ld16(ta + 2 * Cfg::NCOL + Cfg::NS - 8, o8);
      r[4 * 32 * Cfg::RPAD + 32 + lane] = e8[8] + o8[8];
    }
    if (warp == S) {  // accumulator row 32*S (lane 0 of this quadrant): the ones row = column sums of ot (any tap; take n = 1)
      float v[32], m[32];
      const uint32_t ta = tmem + (((uint32_t)S * 32u) << 16);
      tmem_ld32(ta + 1 * Cfg::NCOL, v);
#pragma unroll
      for (int sy = 1; sy < S; ++sy) {
        tmem_ld32(ta + 1 * Cfg::NCOL + sy * 32, m);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += m[c];
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 32; ++c) red[4 * 32 * Cfg::RPAD + c] = v[c];
      }
    }
    fence_before_sync();
    __syncthreads();
    for (int e = tid; e < Cfg::PART; e += NT + 32) {
      float t;
      if (e < 4 * 32 * 32) {
        const int a = (e >> 5) * Cfg::RPAD + (e & 31);
        t = red[a];
#pragma unroll
        for (int w = 1; w < S; ++w) t += red[w * Cfg::RED + a];
      } else if (e < 4 * 32 * 32 + 32) {
        t = red[4 * 32 * Cfg::RPAD + (e - 4 * 32 * 32)];  // column sums of ot (piece-0 region only)
      } else {
        const int a = 4 * 32 * Cfg::RPAD + 32 + (e - 4 * 32 * 32 - 32);  // column sums of ga: add the pieces
        t = red[a];
#pragma unroll
        for (int w = 1; w < S; ++w) t += red[w * Cfg::RED + a];
      }
      out[e] = t;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

static int wg4_split(int precision) { return (precision == VQB_PREC_BF16X3 || precision == VQB_PREC_FP16X2) ? 3 : precision == VQB_PREC_BF16X2 ? 2 : 1; }

static int wgrad4_grid(int B, int Lo, int* tiles_per_b) {
  *tiles_per_b = cdiv(Lo, Wg4Cfg<1>::TK);
  const long total = (long)B * *tiles_per_b;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) num_sms = 148;
  }
  return (int)(total < num_sms ? (total > 0 ? total : 1) : num_sms);
}

// Lo = rows of the short tensor (Conv1D: ceil(L/2) output rows; Conv1DTranspose: L input rows)
size_t wgrad4_tc_workspace_bytes(int B, int Lo) {
  int tpb;
  return (size_t)wgrad4_grid(B, Lo, &tpb) * Wg4Cfg<1>::PART * sizeof(float) + 64;
}

template <int S>
static int launch_wg4(const Wg4Params& p, int grid, cudaStream_t st) {
  using Cfg = Wg4Cfg<S>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(wgrad4_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  VQB_CUDA(launch_pdl(wgrad4_tc_kernel<S>, dim3(grid), dim3(Cfg::NT + 32), (size_t)Cfg::SMEM, st, p));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

// dw [4, 32, 32]; dbias [32] or NULL, = column sums of ot (bias_from_ga = false) or of ga (true)
int wgrad4_tc(int precision, const float* ga, int Lg, const float* ot, int Lo, int B, float* dw, float* dbias, bool bias_from_ga,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = wgrad4_tc_workspace_bytes(B, Lo);
  if (!ws || ws_bytes < need) return set_err(VQB_ERR_WORKSPACE, "tensor-core wgrad workspace: need %zu bytes, got %zu", need, ws_bytes);
  Wg4Params p{};
  p.ga = ga; p.ot = ot; p.partial = (float*)ws;
  p.B = B; p.Lg = Lg; p.Lo = Lo;
  const int grid = wgrad4_grid(B, Lo, &p.tiles_per_b);
  p.total_tiles = B * p.tiles_per_b;
  const int S = wg4_split(precision);
  int rc = S == 3 ? launch_wg4<3>(p, grid, st) : S == 2 ? launch_wg4<2>(p, grid, st) : launch_wg4<1>(p, grid, st);
  if (rc) return rc;
  constexpr int PART = Wg4Cfg<1>::PART;
  reduce_chunks_strided(p.partial, grid, PART, 0, 4 * 32 * 32, dw, st);
  VQB_LAUNCH_CHECK();
  if (dbias) {
    reduce_chunks_strided(p.partial, grid, PART, 4 * 32 * 32 + (bias_from_ga ? 32 : 0), 32, dbias, st);
    VQB_LAUNCH_CHECK();
  }
  return VQB_OK;
}

}  // namespace vqb
