// Weight-streaming linear layer for the decode step (M <= 64 rows): C[M,N] = epilogue(A[M,K] . W[N,K]^T).
//
// At M = 16 the layer is a pure HBM stream of W (2*N*K bytes against 32*N*K flops), so the kernel is built
// around keeping all 148 SMs pulling weights at an equal rate:
//   * swap-AB: 128 rows of W are the UMMA M dimension and the (padded) M activation rows are the UMMA N
//     dimension (16..64), so one tcgen05.mma consumes 4 KB of W for 128 x Mb x 16 MACs -- the tensor pipe
//     never limits the stream (the 128x64 tile of gemm_tcgen05.cu spends 8x more MMA time per weight byte
//     and its single issuing thread caps a CTA at 20-25 GB/s);
//   * stream-K: the (N/128) x (K/64) k-block iterations are cut into gridDim.x equal contiguous ranges, one
//     per SM, so every shape (N = 4096 .. 22016) loads the machine evenly -- no wave quantisation;
//   * a range that covers only part of a tile's K leaves its 128 x Mb fp32 partial in a workspace slot; the
//     LAST contributor to arrive (per-tile counter) sums the slots in slot order -- deterministic, no spin
//     wait, no co-residency requirement -- and applies the epilogue;
//   * programmatic dependent launch: the producer prefetches its first ring of W tiles (weights never depend
//     on the previous kernel) BEFORE griddepcontrol.wait, so the pipeline fill overlaps the tail of whatever
//     ran before; everything that reads activations / residuals or writes happens after the wait.
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue (TMEM lane
// quarter = warp % 4; thread = one W row = one output column, Mb batch values in registers).
// EVO_EPI_GELU_GATE: a tile is 256 packed rows [l1 | l2] (two MMA row groups into adjacent TMEM columns), the
// epilogue thread holds l1[b] and l2[b] of its column: gelu(l1)*l2 is fused (the 128x64 tile needed a second kernel).
#include "common.cuh"
#include "../../include/evo_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

using namespace evo;

namespace {

constexpr int BK = 64;              // bf16 per k-block = one 128-byte swizzle row
constexpr int UK = 16;
constexpr int WROWS = 128;          // W rows per MMA row group (UMMA M)
constexpr int MAXST = 12;           // ring depth upper bound (barrier array size)
constexpr int NTHREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int MAX_MB = 64;

struct SmArgs {
  bf16* C; long long ldc;
  const bf16* bias;
  const bf16* resid; long long ldr;
  int M, Mb;                        // rows of A; rows staged / UMMA N (multiple of 16)
  int n_tiles, KB;                  // tiles of (128*R) W rows; k-blocks per tile
  int n_stages;
  // EPI_HYENA_STEP (fused engine.step_fir + step_iir behind the in-projection): the tile is one head's [x2 | x1 | v] rows
  bf16* fir_state; float* state;    // (B, 3D, 2) bf16, (B, D, 8, 2) fp32: updated in place
  const bf16* fir_w; const bf16* fir_b; const bf16* Dskip;
  const float* poles; const float* residues;
  int D;
  int* counters;                    // [n_tiles], zero between launches
  float* slots;                     // [2 * gridDim.x][R * Mb][128] fp32 partials
  long long* trace;                 // debug: [gridDim.x][16] time stamps (evo_debug_smallm_trace), NULL normally
};

long long* g_trace = nullptr;

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void stamp(const SmArgs& g, int i) { if (g.trace) g.trace[blockIdx.x * 16 + i] = clock64(); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// first k-block iteration of CTA c when `total` iterations are cut into `parts` contiguous ranges.  32-bit on purpose
// (the host checks total * parts < 2^31): 64-bit division is a ~100-instruction subroutine and this sits in the tail.
__device__ __forceinline__ uint32_t range_start(uint32_t c, uint32_t total, uint32_t parts) { return c * total / parts; }
__device__ __forceinline__ int cta_of(uint32_t it, uint32_t total, uint32_t parts) {
  uint32_t c = it * parts / total;
  while (c + 1 < parts && range_start(c + 1, total, parts) <= it) ++c;
  while (c > 0 && range_start(c, total, parts) > it) --c;
  return (int)c;
}

// residual values of 16 batch rows of output column n, fetched BEFORE the accumulators are needed (one round trip
// for all 16 instead of a load -> store chain the compiler may not reorder)
template <int EPI>
__device__ __forceinline__ void load_resid16(const SmArgs& g, long long n, int b0, float (&rv)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    rv[j] = 0.f;
    if constexpr (EPI == EVO_EPI_BIAS_RESID || EPI == EVO_EPI_RESID)
      if (b0 + j < g.M) rv[j] = __bfloat162float(g.resid[(long long)(b0 + j) * g.ldr + n]);
  }
}

// epilogue of 16 batch rows [b0, b0+16) of output column n; v0 (and v1 for the gate) are fp32 sums
template <int EPI>
__device__ __forceinline__ void finalize16(const SmArgs& g, long long n, int b0, float bias, const float (&rv)[16],
                                           const float (&v0)[16], const float (&v1)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int b = b0 + j;
    float o;
    if constexpr (EPI == EVO_EPI_GELU_GATE) {
      o = rbf(gelu_erf(rbf(v0[j]))) * rbf(v1[j]);
    } else {
      o = v0[j];
      if constexpr (EPI == EVO_EPI_BIAS || EPI == EVO_EPI_BIAS_RESID) o += bias;
      if constexpr (EPI == EVO_EPI_BIAS_RESID || EPI == EVO_EPI_RESID) o = rbf(o) + rv[j];
    }
    if (b < g.M) g.C[(long long)b * g.ldc + n] = __float2bfloat16_rn(o);
  }
}

constexpr int EPI_HYENA_STEP = 7;   // EVO_EPI_HYENA_STEP
template <int EPI> struct RowGroups { static constexpr int R = EPI == EVO_EPI_GELU_GATE ? 2 : (EPI == EPI_HYENA_STEP ? 3 : 1); };

// One decode step of the Hyena operator for channel `ch` of head `tile`, 16 batch rows: the arithmetic of hyena_step_kernel
// (hyena.cu; engine.step_fir + step_iir of the reference) on the in-projection's fp32 sums z2/z1/zv (x2, x1, v rows of the head).
// Every rounding point is the one the two-kernel path has: z = bf16(sum + bias) is what EVO_EPI_BIAS would have stored.
__device__ __forceinline__ void hyena_step16(const SmArgs& g, int tile, int row, int b0, const float (&z2)[16], const float (&z1)[16], const float (&zv)[16]) {
  const int D = g.D, ch = tile * WROWS + row;
  const long long C3 = 3LL * D;
  const long long c_base = (long long)tile * 3 * WROWS + row;          // z column of x2; x1 = +128, v = +256
  float w[3][3], fb[3], pb[3];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
    const long long c = c_base + ci * WROWS;
#pragma unroll
    for (int k = 0; k < 3; ++k) w[ci][k] = __bfloat162float(g.fir_w[c * 3 + k]);
    fb[ci] = __bfloat162float(g.fir_b[c]);
    pb[ci] = __bfloat162float(g.bias[c]);
  }
  const float dskip = __bfloat162float(g.Dskip[ch]);
  float2 p[8], r[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    p[s] = reinterpret_cast<const float2*>(g.poles)[(long long)ch * 8 + s];
    r[s] = reinterpret_cast<const float2*>(g.residues)[(long long)ch * 8 + s];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int b = b0 + j;
    if (b >= g.M) continue;
    float f[3];
    const float zin[3] = {z2[j], z1[j], zv[j]};
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const long long c = c_base + ci * WROWS;
      const bf16 un_b = __float2bfloat16_rn(zin[ci] + pb[ci]);
      const float un = __bfloat162float(un_b);
      bf16* fs = g.fir_state + ((long long)b * C3 + c) * 2;
      const bf16 s1_b = fs[1];
      const float s0 = __bfloat162float(fs[0]), s1 = __bfloat162float(s1_b);
      const float t0 = rbf(w[ci][2] * un);
      const float t1 = rbf(rbf(s0 * w[ci][0]) + rbf(s1 * w[ci][1]));
      f[ci] = rbf(rbf(t0 + t1) + fb[ci]);
      fs[0] = s1_b; fs[1] = un_b;
    }
    const float x = rbf(f[1] * f[2]);
    float4* st = reinterpret_cast<float4*>(g.state + ((long long)b * D + ch) * 16);
    float4 sv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) sv[q] = st[q];
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float nr0 = fmaf(p[2 * q].x, sv[q].x, fmaf(-p[2 * q].y, sv[q].y, x));
      const float ni0 = fmaf(p[2 * q].x, sv[q].y, p[2 * q].y * sv[q].x);
      acc = fmaf(r[2 * q].x, nr0, acc); acc = fmaf(-r[2 * q].y, ni0, acc);
      const float nr1 = fmaf(p[2 * q + 1].x, sv[q].z, fmaf(-p[2 * q + 1].y, sv[q].w, x));
      const float ni1 = fmaf(p[2 * q + 1].x, sv[q].w, p[2 * q + 1].y * sv[q].z);
      acc = fmaf(r[2 * q + 1].x, nr1, acc); acc = fmaf(-r[2 * q + 1].y, ni1, acc);
      st[q] = make_float4(nr0, ni0, nr1, ni1);
    }
    g.C[(long long)b * g.ldc + ch] = __float2bfloat16_rn(f[0] * (acc + rbf(dskip * x)));
  }
}

template <int EPI>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_smallm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const SmArgs g) {
  constexpr int R = RowGroups<EPI>::R;                         // MMA row groups per tile
  constexpr int WBOX = R == 2 ? 2 * WROWS : WROWS;             // W rows per TMA box (a box dimension is at most 256)
  constexpr int NBOX = R * WROWS / WBOX;
  constexpr int W_STAGE = R * WROWS * BK * 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int NST = g.n_stages, Mb = g.Mb, KB = g.KB;
  const int A_STAGE = Mb * BK * 2;
  uint8_t* smW = smem;
  uint8_t* smA = smem + NST * W_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smA + NST * A_STAGE);
  uint64_t* empty = full + MAXST;
  uint64_t* tfull = empty + MAXST;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int* flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (g.trace && threadIdx.x == 0) { g.trace[blockIdx.x * 16 + 0] = gtimer(); g.trace[blockIdx.x * 16 + 1] = clock64(); }

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmA); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAXST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  const int acc_cols = R * Mb;                                // fp32 columns of one accumulator
  uint32_t tmem_cols = 32; while ((int)tmem_cols < 2 * acc_cols) tmem_cols *= 2;
  if (warp == 2) { tmem_alloc<1>(tmem_slot, tmem_cols); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(g, 2);                          // setup done

  const uint32_t total = (uint32_t)g.n_tiles * KB, parts = gridDim.x;
  const int it_begin = (int)range_start(blockIdx.x, total, parts), it_end = (int)range_start(blockIdx.x + 1, total, parts);
  const int n_it = it_end - it_begin;
  const uint32_t STAGE_TX = (uint32_t)(W_STAGE + A_STAGE);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      const int tile_b = it_begin / KB, kb_b = it_begin - tile_b * KB;
      const int pre = min(NST, n_it);
      int tile = tile_b, kb = kb_b;
      for (int i = 0; i < pre; ++i) {                         // weights first: they do not depend on the previous kernel
        mbar_arrive_expect_tx(&full[i], STAGE_TX);
#pragma unroll
        for (int bx = 0; bx < NBOX; ++bx) tma_load_2d(smW + i * W_STAGE + bx * (WBOX * BK * 2), &tmW, &full[i], kb * BK, tile * (R * WROWS) + bx * WBOX);
        if (++kb == KB) { kb = 0; ++tile; }
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      tile = tile_b; kb = kb_b;
      for (int i = 0; i < pre; ++i) {
        tma_load_2d(smA + i * A_STAGE, &tmA, &full[i], kb * BK, 0);
        if (++kb == KB) { kb = 0; ++tile; }
      }
      int stage = pre == NST ? 0 : pre; uint32_t phase = pre == NST ? 1 : 0;
      for (int i = pre; i < n_it; ++i) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], STAGE_TX);
#pragma unroll
        for (int bx = 0; bx < NBOX; ++bx) tma_load_2d(smW + stage * W_STAGE + bx * (WBOX * BK * 2), &tmW, &full[stage], kb * BK, tile * (R * WROWS) + bx * WBOX);
        tma_load_2d(smA + stage * A_STAGE, &tmA, &full[stage], kb * BK, 0);
        if (++kb == KB) { kb = 0; ++tile; }
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(WROWS, Mb);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int it = it_begin;
      while (it < it_end) {
        const int tile = it / KB, k0 = it - tile * KB;
        const int k1 = min(KB, k0 + (it_end - it));
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&full[stage], phase);
          if (it == it_begin && kb == k0) stamp(g, 3);        // first stage landed
          tc_fence_after();
          const uint64_t bd = umma_desc_k_sw128(smem_u32(smA + stage * A_STAGE));
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const uint64_t ad = umma_desc_k_sw128(smem_u32(smW + stage * W_STAGE + r * (WROWS * BK * 2)));
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_ss<1>(d_tmem + (uint32_t)(r * Mb), ad + (uint64_t)(k * UK * 2 / 16), bd + (uint64_t)(k * UK * 2 / 16), idesc, (kb != k0 || k != 0));
          }
          umma_commit(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
        it += k1 - k0;
      }
      stamp(g, 4);                                            // all MMAs issued
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------ epilogue / stream-K fix-up
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int q = warp - EPI_WARP0;
    const int row = q * 32 + lane;                            // W row inside the row group == output column inside the tile
    const int slot_floats = acc_cols * WROWS;
    int acc = 0; uint32_t acc_phase = 0;
    int it = it_begin;
    const int my_first_tile = it_begin / KB;
    while (it < it_end) {
      const int tile = it / KB, k0 = it - tile * KB;
      const int k1 = min(KB, k0 + (it_end - it));
      const bool whole = k0 == 0 && k1 == KB;
      const long long n = (long long)tile * WROWS + row;      // output column (gate: of the N/2-wide output)
      float bias = 0.f;
      if constexpr (EPI == EVO_EPI_BIAS || EPI == EVO_EPI_BIAS_RESID) bias = __bfloat162float(g.bias[n]);
      mbar_wait(&tfull[acc], acc_phase);
      if (threadIdx.x == EPI_WARP0 * 32) stamp(g, it + (k1 - k0) >= it_end ? 5 : 9);   // (last) accumulator complete
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_cols);
      if (whole) {
        for (int c = 0; c < Mb; c += 16) {
          uint32_t r0[16], r1[16], r2[16];
          float rv[16];
          tmem_ld_32x16(t0 + c, r0);
          if constexpr (R >= 2) tmem_ld_32x16(t0 + Mb + c, r1);
          if constexpr (R == 3) tmem_ld_32x16(t0 + 2 * Mb + c, r2);
          load_resid16<EPI>(g, n, c, rv);
          tmem_ld_wait();
          float v0[16], v1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { v0[j] = __uint_as_float(r0[j]); v1[j] = R >= 2 ? __uint_as_float(r1[j]) : 0.f; }
          if constexpr (EPI == EPI_HYENA_STEP) {
            float v2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v2[j] = __uint_as_float(r2[j]);
            hyena_step16(g, tile, row, c, v0, v1, v2);
          } else finalize16<EPI>(g, n, c, bias, rv, v0, v1);
        }
      } else {
        float* slot = g.slots + (size_t)(2 * blockIdx.x + (tile == my_first_tile ? 0 : 1)) * slot_floats;
        for (int c = 0; c < acc_cols; c += 16) {
          uint32_t r0[16];
          tmem_ld_32x16(t0 + c, r0);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) slot[(c + j) * WROWS + row] = __uint_as_float(r0[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);               // the accumulator is free once it has been read
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
      if (!whole) {
        // publication: the CTA barrier orders every thread's partial stores before thread 0's acq_rel atomic (release
        // is cumulative), whose acquire half orders the finisher's slot reads after the other contributors' releases
        const int c_first = cta_of((uint32_t)tile * KB, total, parts), c_last = cta_of((uint32_t)(tile + 1) * KB - 1, total, parts);
        named_bar(1, 128);
        if (threadIdx.x == EPI_WARP0 * 32) {
          int old;
          asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(g.counters + tile) : "memory");
          *flag = old;
        }
        named_bar(1, 128);
        const bool last = *flag == c_last - c_first;
        if (threadIdx.x == EPI_WARP0 * 32) stamp(g, 6);       // partial published
        if (last) {
          for (int c = 0; c < Mb; c += 16) {
            float v0[16], v1[16], v2[16], rv[16];
            load_resid16<EPI>(g, n, c, rv);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v0[j] = 0.f; v1[j] = 0.f; v2[j] = 0.f; }
            // slots are summed in contributor order (deterministic); the loads of FX contributors are in flight together
            constexpr int FX = R == 3 ? 2 : (R == 2 ? 4 : 8);
            for (int cb = c_first; cb <= c_last; cb += FX) {
              float t0v[FX][16], t1v[FX][16], t2v[FX][16];
#pragma unroll
              for (int u = 0; u < FX; ++u) {
                const int cc = min(cb + u, c_last);
                const int first_tile_cc = (int)(range_start((uint32_t)cc, total, parts) / (uint32_t)KB);
                const float* s = g.slots + (size_t)(2 * cc + (first_tile_cc == tile ? 0 : 1)) * slot_floats;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  t0v[u][j] = __ldcg(s + (c + j) * WROWS + row);
                  if constexpr (R >= 2) t1v[u][j] = __ldcg(s + (Mb + c + j) * WROWS + row);
                  if constexpr (R == 3) t2v[u][j] = __ldcg(s + (2 * Mb + c + j) * WROWS + row);
                }
              }
#pragma unroll
              for (int u = 0; u < FX; ++u) {
                if (cb + u <= c_last) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) { v0[j] += t0v[u][j]; if constexpr (R >= 2) v1[j] += t1v[u][j]; if constexpr (R == 3) v2[j] += t2v[u][j]; }
                }
              }
            }
            if constexpr (EPI == EPI_HYENA_STEP) hyena_step16(g, tile, row, c, v0, v1, v2);
            else finalize16<EPI>(g, n, c, bias, rv, v0, v1);
          }
          if (threadIdx.x == EPI_WARP0 * 32) g.counters[tile] = 0;   // self-cleaning for the next launch
        }
      }
      it += k1 - k0;
    }
  }

  if (threadIdx.x == EPI_WARP0 * 32) stamp(g, 7);             // epilogue done
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, tmem_cols);
  if (g.trace && threadIdx.x == 0) { g.trace[blockIdx.x * 16 + 8] = clock64(); g.trace[blockIdx.x * 16 + 15] = gtimer(); }
}

int smem_budget_bytes() {
  static int kb = 0;
  if (!kb) { const char* e = getenv("EVO_B200_SMALLM_SMEM_KB"); kb = e ? atoi(e) : 104; kb = std::max(40, std::min(kb, 224)); }
  return kb * 1024;
}

template <int EPI>
int launch(const evo_gemm_smallm_params* p, cudaStream_t st) {
  constexpr int R = RowGroups<EPI>::R;
  const int Mb = (int)((p->M + 15) / 16 * 16);
  CUtensorMap tmW, tmA;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmW, p->W, (uint64_t)p->K, (uint64_t)p->N, (uint64_t)p->K * 2, BK, R == 2 ? 2 * WROWS : WROWS, true))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmA, p->A, (uint64_t)p->K, (uint64_t)p->M, (uint64_t)p->lda * 2, BK, (uint32_t)Mb, true))) return rc;
  SmArgs g;
  g.C = (bf16*)p->C; g.ldc = p->ldc; g.bias = (const bf16*)p->bias; g.resid = (const bf16*)p->residual; g.ldr = p->ldr;
  g.M = (int)p->M; g.Mb = Mb;
  g.fir_state = (bf16*)p->fir_state; g.state = p->state; g.fir_w = (const bf16*)p->fir_w; g.fir_b = (const bf16*)p->fir_b; g.Dskip = (const bf16*)p->Dskip;
  g.poles = p->poles; g.residues = p->residues; g.D = (int)(p->N / 3);
  g.n_tiles = (int)(p->N / (R * WROWS));
  g.KB = (int)(p->K / BK);
  const int stage_bytes = R * WROWS * BK * 2 + Mb * BK * 2;
  const int tail = 2 * MAXST * 8 + 4 * 8 + 16;
  // three row groups stage 48 KB of W per k-block: give the ring 160 KB (3 stages) instead of the default budget (2)
  const int budget = R == 3 ? std::max(smem_budget_bytes(), 160 * 1024) : smem_budget_bytes();
  g.n_stages = std::max(2, std::min(MAXST, (budget - tail) / stage_bytes));
  const int smem_bytes = g.n_stages * stage_bytes + tail;
  const long long total = (long long)g.n_tiles * g.KB;
  EVO_REQUIRE(total * device_sm_count() < (1LL << 31), "evo_gemm_smallm: N * K too large (%lld k-block iterations)", total);
  // one range per SM, but at least 8 k-blocks per range (bounds the contributors per tile for tiny layers)
  const int grid = (int)std::max<long long>(1, std::min<long long>(device_sm_count(), total / 8));
  g.trace = g_trace;
  const size_t need = evo_gemm_smallm_workspace(p->M, p->N, p->K, p->epilogue);
  EVO_REQUIRE(p->workspace != nullptr && p->workspace_bytes >= need, "evo_gemm_smallm: workspace too small (%zu < %zu)", p->workspace_bytes, need);
  g.counters = (int*)p->workspace;
  g.slots = (float*)((uint8_t*)p->workspace + 16384);
  static unsigned long long attr_done = 0;
  auto kern = gemm_smallm_kernel<EPI>;
  { int rc_ = ensure_dyn_smem(kern, 224 * 1024, attr_done); if (rc_) return rc_; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_level() >= 1) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  EVO_CUDA(cudaLaunchKernelEx(&cfg, kern, tmW, tmA, g));
  return check_launch("evo_gemm_smallm");
}

}  // namespace

#ifdef EVO_SMALLM_TRACE   // in-kernel time stamps: experiment builds only (nvcc -DEVO_SMALLM_TRACE), not part of the shipped ABI
extern "C" void evo_debug_smallm_trace(void* buf) { g_trace = (long long*)buf; }
#endif

extern "C" size_t evo_gemm_smallm_workspace(int64_t M, int64_t N, int64_t K, int epilogue) {
  (void)K;
  const int R = epilogue == EVO_EPI_GELU_GATE ? 2 : (epilogue == EPI_HYENA_STEP ? 3 : 1);
  const size_t Mb = (size_t)((M + 15) / 16 * 16);
  if (N / (R * WROWS) > 4096) return 0;
  return 16384 + (size_t)2 * device_sm_count() * R * Mb * WROWS * sizeof(float);
}

extern "C" int evo_gemm_smallm(const evo_gemm_smallm_params* p, void* stream) {
  EVO_REQUIRE(p->M >= 0 && p->M <= MAX_MB, "evo_gemm_smallm: M (%lld) must be in [0, %d]", (long long)p->M, MAX_MB);
  EVO_REQUIRE(p->N > 0 && p->K > 0 && p->K % BK == 0, "evo_gemm_smallm: K (%lld) must be a positive multiple of %d", (long long)p->K, BK);
  EVO_REQUIRE(p->N % 256 == 0 && p->N / WROWS <= 4096, "evo_gemm_smallm: N (%lld) must be a multiple of 256 (and <= 524288)", (long long)p->N);
  EVO_REQUIRE(p->lda % 8 == 0, "evo_gemm_smallm: lda must be a multiple of 8 elements");
  EVO_REQUIRE(((uintptr_t)p->A % 16) == 0 && ((uintptr_t)p->W % 16) == 0, "evo_gemm_smallm: A and W must be 16-byte aligned");
  static_assert(EPI_HYENA_STEP == EVO_EPI_HYENA_STEP, "enum drift");
  if (p->epilogue == EVO_EPI_BIAS || p->epilogue == EVO_EPI_BIAS_RESID) EVO_REQUIRE(p->bias != nullptr, "evo_gemm_smallm: bias epilogue without bias");
  if (p->epilogue == EVO_EPI_RESID || p->epilogue == EVO_EPI_BIAS_RESID) EVO_REQUIRE(p->residual != nullptr, "evo_gemm_smallm: residual epilogue without residual");
  if (p->M == 0) return 0;
  switch (p->epilogue) {
    case EVO_EPI_NONE: return launch<EVO_EPI_NONE>(p, (cudaStream_t)stream);
    case EVO_EPI_BIAS: return launch<EVO_EPI_BIAS>(p, (cudaStream_t)stream);
    case EVO_EPI_BIAS_RESID: return launch<EVO_EPI_BIAS_RESID>(p, (cudaStream_t)stream);
    case EVO_EPI_RESID: return launch<EVO_EPI_RESID>(p, (cudaStream_t)stream);
    case EVO_EPI_GELU_GATE: return launch<EVO_EPI_GELU_GATE>(p, (cudaStream_t)stream);
    case EPI_HYENA_STEP:
      EVO_REQUIRE(p->bias && p->fir_state && p->state && p->fir_w && p->fir_b && p->Dskip && p->poles && p->residues, "evo_gemm_smallm: EVO_EPI_HYENA_STEP needs bias, states and filter parameters");
      EVO_REQUIRE(p->N % (3 * WROWS) == 0, "evo_gemm_smallm: EVO_EPI_HYENA_STEP needs N = 3 * heads * 128");
      return launch<EPI_HYENA_STEP>(p, (cudaStream_t)stream);
  }
  set_error("evo_gemm_smallm: unknown epilogue %d", p->epilogue);
  return -1;
}
