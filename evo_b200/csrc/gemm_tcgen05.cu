// Tensor-core linear layers for sm_100a: C[M,N] = epilogue(A[M,K] . W[N,K]^T).
//
// Persistent, warp-specialised tcgen05 kernel:
//   warp 0   TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1   MMA issuer     (one thread issues tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warp 2   TMEM allocator
//   warps 4-7 epilogue      (tcgen05.ld -> fused bias / residual / GELU-gate -> bf16 -> HBM)
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i
// overlaps the mainloop of tile i+1.
//
// CG = 1: one CTA per 128 x 256 output tile (UMMA 128x256x16), 4-stage ring.
// CG = 2: a CTA pair (cluster 2x1) per 256 x 256 tile (UMMA 256x256x16, cta_group::2): each
//         CTA stages its own 128 rows of A and 128 rows of W, halving per-SM operand traffic;
//         6-stage ring.  The leader CTA issues the MMAs and multicasts the completions.
//
// Epilogues reproduce the reference's bf16 rounding points (nn.Linear output is rounded to
// bf16 before the residual add / the gate): see include/evo_b200.h.
#include "common.cuh"
#include "../../include/evo_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

using namespace evo;

namespace {

constexpr int BM = 128;            // rows of C per CTA
constexpr int BN_BIG = 256;        // columns of C per tile (UMMA N) for the throughput tiles
constexpr int BN_SMALL = 64;       // weight-streaming tiles for small M (decode): N/64 CTAs keep every SM pulling HBM
constexpr int BK = 64;             // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UK = 16;             // UMMA K for 16-bit inputs
constexpr int NTHREADS = 256;
constexpr int EPI_WARP0 = 4;

template <int CG, int BN> struct Cfg {
  static constexpr int B_ROWS = BN / CG;                       // W rows staged per CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // Small-M tile: A is staged COMPACT (only a_rows <= 128 rows per k-block; the MMA still reads a 128-row tile whose
  // remaining rows alias later stages / padding and only feed accumulator rows that are never stored), which leaves
  // room for a 16-deep ring of W tiles: bytes in flight, not tensor throughput, bound the weight-streaming GEMM.
  // The weight-streaming kernel is bound by the serial wait -> 4 tiny MMAs -> commit round of its single issuing
  // thread (~250 cycles per 8 KB k-block, measured 20-25 GB/s per CTA), not by bytes in flight: two CTAs per SM
  // (<= 110 KB of shared memory each, 128 TMEM columns each) double the number of issuing threads.
  static constexpr int STAGES = BN == BN_SMALL ? 8 : (CG == 1 ? 4 : 6);
  static constexpr int SMEM_BYTES = BN == BN_SMALL ? 110 * 1024 : STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int CTAS_PER_SM = BN == BN_SMALL ? 2 : 1;
};

struct GemmArgs {
  bf16* C; long long ldc;
  const bf16* bias;
  const bf16* resid; long long ldr;
  long long M, N, K;
  int m_blocks, n_blocks, group_m;
  int b_tiled;           // small-M tile only: W is tile-major [N/64][K/64][64][64] (contiguous 8 KB tiles: every DRAM page opened is fully used)
  int a_rows, n_stages, ksub;  // small-M tile only: rows of A staged per k-block, ring depth, 64-column k-blocks per ring stage
  // peer-scattered output (sequence-parallel Ulysses re-shard fused into the Wqkv epilogue): column n of row m goes to
  // c_peer[(n % peer_period) / peer_inner] at [(peer_row0 + m) * ldc + (n / peer_period) * peer_inner + n % peer_inner]
  bf16* c_peer[8]; int n_peers; int peer_period, peer_inner; long long peer_row0;
  const bf16* rope_cos; const bf16* rope_sin;   // EVO_EPI_BIAS_ROPE: (positions, 64) bf16 tables, row l = position of token row l
  long long rope_L, rope_cols;                  //   rows repeat with period rope_L (tokens per sequence); columns < rope_cols (q and k) are rotated
  const long long* targets;   // EPI_LSE only: (M) target token per row (-1: none)
  float4* part;               // EPI_LSE only: (M, n_blocks) per-row partial statistics {max, sum e^(x-max), sum e^(x-max) x, target logit}
  int l2_hints;          // TMA loads carry L2 eviction priorities (resident slab evict_last, streaming operand evict_first)
  const uint16_t* die_tab;   // die-aware rasterisation (CG == 2): tab[smid >> 1] = die | slot << 1, nullptr = off
  unsigned* die_claim;       // one word per (die, slot), zeroed before the launch: a pair owns the slot it sets first
  int m_split, die_pairs0, die_pairs1;   // die 0 owns row-blocks [0, m_split) with die_pairs0 CTA pairs, die 1 the rest
  int die_collide;           // test hook: every pair asks for the same slot first
  int skew;              // experiment (EVO_B200_GEMM_SKEW): the producer of tile slot t starts t * skew cycles late
  int raster_n;          // 0: groups of `group_m` row-blocks sweep all of N (A stays in L2); 1: groups of `group_m` column-blocks sweep all of M (W stays in L2)
};

constexpr int EPI_LSE = 5;     // internal: scoring epilogue (evo_unembed_score); no C is written (EVO_EPI_BIAS_ROPE is 6)

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// grouped rasterisation: GROUP_M row-blocks share the sweep over n so their A tiles stay in L2
// (over the row-blocks [m_lo, m_lo + m_cnt): the whole matrix, or the share of one die)
__device__ __forceinline__ void tile_coords(int tile, const GemmArgs& g, int m_lo, int m_cnt, int& m_blk, int& n_blk) {
  const int major = g.raster_n ? g.n_blocks : m_cnt;     // the grouped dimension
  const int minor = g.raster_n ? m_cnt : g.n_blocks;     // the swept dimension
  int per_group = g.group_m * minor;
  int grp = tile / per_group;
  int first = grp * g.group_m;
  int gsz = min(g.group_m, major - first);
  int in = tile - grp * per_group;
  int a = first + in % gsz, b = in / gsz;
  m_blk = m_lo + (g.raster_n ? b : a);
  n_blk = g.raster_n ? a : b;
}

// address of C[row, col] (col = first column of a 32-wide chunk; a chunk never straddles a head)
__device__ __forceinline__ bf16* c_ptr(const GemmArgs& g, long long row, int col) {
  if (g.n_peers == 0) return g.C + row * g.ldc + col;
  const int t = col / g.peer_period, rem = col - t * g.peer_period;
  const int p = rem / g.peer_inner;
  return g.c_peer[p] + (g.peer_row0 + row) * g.ldc + t * g.peer_inner + (rem - p * g.peer_inner);
}

template <int EPI>
__device__ __forceinline__ void store_chunk(const GemmArgs& g, long long row, int col, const uint32_t (&acc)[32], const uint32_t (&acc2)[32]) {
  // acc: 32 consecutive fp32 accumulator columns of this thread's row (as raw bits)
  uint32_t outw[16];
  if constexpr (EPI == EVO_EPI_GELU_GATE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a0 = rbf(__uint_as_float(acc[2 * j])), a1 = rbf(__uint_as_float(acc[2 * j + 1]));
      float b0 = rbf(__uint_as_float(acc2[2 * j])), b1 = rbf(__uint_as_float(acc2[2 * j + 1]));
      outw[j] = pack_bf16(rbf(gelu_erf(a0)) * b0, rbf(gelu_erf(a1)) * b1);
    }
  } else {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
    if constexpr (EPI == EVO_EPI_BIAS || EPI == EVO_EPI_BIAS_RESID) {
      const uint4* bp = reinterpret_cast<const uint4*>(g.bias + col);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 bv = __ldg(bp + q);
        const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bv);
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[q * 8 + 2 * j] += bf_lo(bw[j]); v[q * 8 + 2 * j + 1] += bf_hi(bw[j]); }
      }
    }
    if constexpr (EPI == EVO_EPI_BIAS_RESID || EPI == EVO_EPI_RESID) {
      const uint4* rp = reinterpret_cast<const uint4*>(g.resid + row * g.ldr + col);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 rv = __ldg(rp + q);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[q * 8 + 2 * j] = rbf(v[q * 8 + 2 * j]) + bf_lo(rw[j]);
          v[q * 8 + 2 * j + 1] = rbf(v[q * 8 + 2 * j + 1]) + bf_hi(rw[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) outw[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
  }
  uint4* dst = reinterpret_cast<uint4*>(c_ptr(g, row, col));
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = make_uint4(outw[4 * q], outw[4 * q + 1], outw[4 * q + 2], outw[4 * q + 3]);
}

// EVO_EPI_BIAS_ROPE: columns [col, col+32) and [col+64, col+96) of one 128-wide head: x = bf16(acc + bias) (the reference's qkv
// tensor), then the NeoX rotation with the token's cos/sin row in fp32, one rounding on store -- the arithmetic of
// rotary_qk_kernel (elementwise.cu), which is flash_attn's apply_rotary (layers/rotary.py:382-416, mha.py:648).
__device__ __forceinline__ void store_rope_pair(const GemmArgs& g, long long row, int col, long long l, int c, bool rotate,
                                                const uint32_t (&acc1)[32], const uint32_t (&acc2)[32]) {
  float v1[32], v2[32];
  const uint4* b1 = reinterpret_cast<const uint4*>(g.bias + col);
  const uint4* b2 = reinterpret_cast<const uint4*>(g.bias + col + 64);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 x = __ldg(b1 + q), y = __ldg(b2 + q);
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&x);
    const uint32_t* yw = reinterpret_cast<const uint32_t*>(&y);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v1[q * 8 + 2 * j] = rbf(__uint_as_float(acc1[q * 8 + 2 * j]) + bf_lo(xw[j])); v1[q * 8 + 2 * j + 1] = rbf(__uint_as_float(acc1[q * 8 + 2 * j + 1]) + bf_hi(xw[j]));
      v2[q * 8 + 2 * j] = rbf(__uint_as_float(acc2[q * 8 + 2 * j]) + bf_lo(yw[j])); v2[q * 8 + 2 * j + 1] = rbf(__uint_as_float(acc2[q * 8 + 2 * j + 1]) + bf_hi(yw[j]));
    }
  }
  uint32_t o1[16], o2[16];
  if (rotate) {
    const uint4* cp = reinterpret_cast<const uint4*>(g.rope_cos + l * 64 + c);
    const uint4* sp = reinterpret_cast<const uint4*>(g.rope_sin + l * 64 + c);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 cv = __ldg(cp + q), sv = __ldg(sp + q);
      const uint32_t* cw = reinterpret_cast<const uint32_t*>(&cv);
      const uint32_t* sw = reinterpret_cast<const uint32_t*>(&sv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = q * 8 + 2 * j;
        const float cl = bf_lo(cw[j]), ch = bf_hi(cw[j]), sl = bf_lo(sw[j]), sh = bf_hi(sw[j]);
        o1[q * 4 + j] = pack_bf16(v1[e] * cl - v2[e] * sl, v1[e + 1] * ch - v2[e + 1] * sh);
        o2[q * 4 + j] = pack_bf16(v1[e] * sl + v2[e] * cl, v1[e + 1] * sh + v2[e + 1] * ch);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) { o1[j] = pack_bf16(v1[2 * j], v1[2 * j + 1]); o2[j] = pack_bf16(v2[2 * j], v2[2 * j + 1]); }
  }
  uint4* d1 = reinterpret_cast<uint4*>(c_ptr(g, row, col));
  uint4* d2 = reinterpret_cast<uint4*>(c_ptr(g, row, col + 64));
#pragma unroll
  for (int q = 0; q < 4; ++q) { d1[q] = make_uint4(o1[4 * q], o1[4 * q + 1], o1[4 * q + 2], o1[4 * q + 3]); d2[q] = make_uint4(o2[4 * q], o2[4 * q + 1], o2[4 * q + 2], o2[4 * q + 3]); }
}

template <int CG, int EPI, int BN>
__global__ void __launch_bounds__(NTHREADS, (BN == BN_SMALL ? 2 : 1))
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using C_ = Cfg<CG, BN>;
  static_assert(EPI != EVO_EPI_GELU_GATE || BN == BN_BIG, "the gate epilogue needs [l1 | l2] halves of a 256-column tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr bool SMALL = BN == BN_SMALL;
  const int NST = SMALL ? g.n_stages : C_::STAGES;                    // ring depth
  const int A_STRIDE = SMALL ? g.a_rows * BK * 2 : C_::A_BYTES;       // bytes of A per stage
  const int KSUB = SMALL ? g.ksub : 1;                                // k-blocks per ring stage (one barrier round)
  const uint32_t STAGE_TX = (uint32_t)KSUB * ((uint32_t)A_STRIDE + C_::B_BYTES);
  uint8_t* smA = smem;
  uint8_t* smB = smem + (SMALL ? ((NST * KSUB * A_STRIDE + (C_::A_BYTES - A_STRIDE) + 1023) & ~1023) : C_::STAGES * C_::A_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smB + NST * KSUB * C_::B_BYTES);
  uint64_t* empty = full + C_::STAGES;
  uint64_t* tfull = empty + C_::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* walk_slot = tmem_slot + 2;          // pair leader: die | slot << 1 of this pair (die-aware rasterisation)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if constexpr (CG == 2) {
    if (g.die_tab != nullptr && warp == 3 && lane == 0 && leader) {
      // The slot this TPC would like (its rank among the TPCs of its die), then a claim: if the slot is taken -- two pairs ran on
      // one TPC one after the other because something else held an SM when the grid started -- take the next free one.  As many
      // slots as pairs, so every slot ends up owned exactly once whatever the placement was.
      uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      const uint32_t want = g.die_collide ? 0u : g.die_tab[sm >> 1];
      const int total = g.die_pairs0 + g.die_pairs1;
      int idx = (int)(want >> 1) + ((want & 1u) ? g.die_pairs0 : 0);
      for (int k = 0; k < total && atomicCAS(&g.die_claim[idx], 0u, 1u) != 0u; ++k) idx = idx + 1 == total ? 0 : idx + 1;
      const uint32_t die = idx >= g.die_pairs0 ? 1u : 0u;
      *walk_slot = die | (uint32_t)(die ? idx - g.die_pairs0 : idx) << 1;
    }
  }
  if (warp == 1 && lane == 0) {
    // full: one arrival (the leader's arrive.expect_tx); the byte count covers both CTAs' loads
    for (int i = 0; i < C_::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * CG); }
    fence_barrier_init();
  }
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulator buffers (power of two >= 32)
  if (warp == 2) { tmem_alloc<CG>(tmem_slot, TMEM_COLS); tmem_relinquish<CG>(); }
  tc_fence_before();
  if constexpr (CG == 2) { cluster_sync_all(); __syncthreads(); }   // (the CTA barrier only makes the hand-off visible to racecheck)
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Tile walk of this CTA (pair): tiles w_first, w_first + w_step, ... < n_tiles of the row-blocks [m_lo, m_lo + m_cnt).
  // Die-aware (CG == 2, full grid): each die works on its own contiguous share of the row-blocks, so the streamed operand's
  // tiles are fetched by one die's L2 only; the pair's (die, slot) comes from the leader's %smid and is read by BOTH CTAs from the
  // leader's shared memory, so the two can never disagree about the walk.
  int w_first = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // m_blocks counts (BM*CG)-row blocks
  int w_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  int m_lo = 0, m_cnt = g.m_blocks;
  if constexpr (CG == 2) {
    if (g.die_tab != nullptr) {
      const uint32_t v = ld_shared_cluster_u32(walk_slot, 0);
      const int die = (int)(v & 1u);
      w_first = (int)(v >> 1);
      w_step = die ? g.die_pairs1 : g.die_pairs0;
      m_lo = die ? g.m_split : 0;
      m_cnt = die ? g.m_blocks - g.m_split : g.m_split;
    }
  }
  const int tile0 = w_first, tile_step = w_step;
  const int n_tiles = m_cnt * g.n_blocks;
  const int nkb = (int)(g.K / (BK * KSUB));

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t full_leader_mask = 0xFEFFFFFFu;   // shared::cluster address of the pair leader's copy
      // L2 eviction priorities (experiment, EVO_B200_GEMM_L2_HINTS; default 0 = plain loads).  1: resident slab evict_last, streaming
      // operand evict_first -- measured WORSE on B200 (profiles/r02_gemm_l2_hints_call11.txt: DRAM bytes 5.5 -> 10.6 GB for the
      // projection, 89.2 -> 84.0 k nt/s): an evict_first tile is dropped before the other CTAs of the wave have fetched it.
      // 2: only the slab is marked (evict_last), the stream keeps the default policy.
      const uint64_t keep = l2_policy_evict_last(), first = g.l2_hints == 1 ? l2_policy_evict_first() : l2_policy_evict_normal();
      const uint64_t pol_a = g.raster_n ? first : keep;
      const uint64_t pol_b = g.raster_n ? keep : first;
      if (g.skew > 0) {      // de-phase the CTAs that share operand tiles: the first of them misses, the others find the tile in L2
        const long long t0 = clock64(), wait = (long long)tile0 * g.skew;
        while (clock64() - t0 < wait) __nanosleep(64);
      }
      for (int tile = tile0; tile < n_tiles; tile += tile_step) {
        int m_blk, n_blk; tile_coords(tile, g, m_lo, m_cnt, m_blk, n_blk);
        const int a_row = (m_blk * CG + (int)cta_rank) * BM;
        const int b_row = n_blk * BN + (int)cta_rank * C_::B_ROWS;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(&full[stage], STAGE_TX);
            for (int u = 0; u < KSUB; ++u) {
              if (!SMALL && g.l2_hints) {
                tma_load_2d_hint(smA + (stage * KSUB + u) * A_STRIDE, &tmA, &full[stage], (kb * KSUB + u) * BK, a_row, pol_a);
                tma_load_2d_hint(smB + (stage * KSUB + u) * C_::B_BYTES, &tmB, &full[stage], (kb * KSUB + u) * BK, b_row, pol_b);
                continue;
              }
              tma_load_2d(smA + (stage * KSUB + u) * A_STRIDE, &tmA, &full[stage], (kb * KSUB + u) * BK, a_row);
              if (SMALL && g.b_tiled) tma_load_4d(smB + (stage * KSUB + u) * C_::B_BYTES, &tmB, &full[stage], 0, 0, kb * KSUB + u, n_blk);
              else tma_load_2d(smB + (stage * KSUB + u) * C_::B_BYTES, &tmB, &full[stage], (kb * KSUB + u) * BK, b_row);
            }
          } else {
            // no remote arrive from the peer: a release.cluster arrive per stage costs more than the
            // 512-cycle k-block budget; the peer's bytes are already counted in the leader's expect_tx
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C_::STAGE_BYTES);
            const uint32_t bar = smem_u32(&full[stage]) & full_leader_mask;
            if (g.l2_hints) {
              tma_load_2d_2sm_hint(smA + stage * C_::A_BYTES, &tmA, bar, kb * BK, a_row, pol_a);
              tma_load_2d_2sm_hint(smB + stage * C_::B_BYTES, &tmB, bar, kb * BK, b_row, pol_b);
            } else {
              tma_load_2d_2sm(smA + stage * C_::A_BYTES, &tmA, bar, kb * BK, a_row);
              tma_load_2d_2sm(smB + stage * C_::B_BYTES, &tmB, bar, kb * BK, b_row);
            }
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (pair leader only)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          for (int u = 0; u < KSUB; ++u) {
            const uint64_t ad = umma_desc_k_sw128(smem_u32(smA + (stage * KSUB + u) * A_STRIDE));
            const uint64_t bd = umma_desc_k_sw128(smem_u32(smB + (stage * KSUB + u) * C_::B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)   // +32 B per UMMA_K inside the 128B swizzle row
              umma_ss<CG>(d_tmem, ad + (uint64_t)(k * UK * 2 / 16), bd + (uint64_t)(k * UK * 2 / 16), idesc, (kb | u | k) != 0);
          }
          if constexpr (CG == 1) umma_commit(&empty[stage]); else umma_commit_2sm(&empty[stage], 0b11);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 1) umma_commit(&tfull[acc]); else umma_commit_2sm(&tfull[acc], 0b11);
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------ epilogue: TMEM -> registers -> HBM
    const int q = warp - EPI_WARP0;                      // == warp % 4: the TMEM lane quarter this warp may read
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      int m_blk, n_blk; tile_coords(tile, g, m_lo, m_cnt, m_blk, n_blk);
      const long long row = (long long)(m_blk * CG + (int)cta_rank) * BM + q * 32 + lane;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * BN;
      if constexpr (EPI == EPI_LSE) {
        // scoring epilogue: the 256 logits of this row's half never leave the SM.  Online max / sum-exp / sum-exp-times-logit
        // over the bf16-rounded logits (rp: the reference's logits tensor is bf16), plus the target's logit if it is in this half.
        const long long tgt = row < g.M && g.targets ? g.targets[row] - (long long)n_blk * BN : -1;
        float m_run = -INFINITY, s_run = 0.f, w_run = 0.f, t_logit = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t r1[32];
          tmem_ld_32x32(t0 + c, r1);
          tmem_ld_wait();
          float v[32], cmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = rbf(__uint_as_float(r1[j])); cmax = fmaxf(cmax, v[j]); }
          if (cmax > m_run) { const float sc = __expf(m_run - cmax); s_run *= sc; w_run *= sc; m_run = cmax; }
          const int tj = (int)(tgt - c);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = __expf(v[j] - m_run);
            s_run += e; w_run = fmaf(e, v[j], w_run);
            if (j == tj) t_logit = v[j];
          }
        }
        if (row < g.M) g.part[row * g.n_blocks + n_blk] = make_float4(m_run, s_run, w_run, t_logit);
      } else if constexpr (EPI == EVO_EPI_BIAS_ROPE && BN == BN_BIG) {
        const bool rotate = (long long)n_blk * BN < g.rope_cols;
        const long long l = row < g.M ? row % g.rope_L : 0;
#pragma unroll 1
        for (int hc = 0; hc < BN; hc += 128) {
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t r1[32], r2[32];
            tmem_ld_32x32(t0 + hc + c, r1);
            tmem_ld_32x32(t0 + hc + 64 + c, r2);
            tmem_ld_wait();
            if (row < g.M) store_rope_pair(g, row, n_blk * BN + hc + c, l, c, rotate, r1, r2);
          }
        }
      } else if constexpr (EPI == EVO_EPI_GELU_GATE && BN == BN_BIG) {
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 32) {
          uint32_t r1[32], r2[32];
          tmem_ld_32x32(t0 + c, r1);
          tmem_ld_32x32(t0 + BN / 2 + c, r2);
          tmem_ld_wait();
          if (row < g.M) store_chunk<EPI>(g, row, n_blk * (BN / 2) + c, r1, r2);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t r1[32];
          tmem_ld_32x32(t0 + c, r1);
          tmem_ld_wait();
          if (row < g.M) store_chunk<EPI>(g, row, n_blk * BN + c, r1, r1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CG == 1) mbar_arrive(&tempty[acc]); else mbar_arrive_cluster_relaxed(&tempty[acc], 0); }
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }

  __syncwarp();            // single-lane roles rejoin their warp before the aligned barriers
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

template <int CG, int EPI, int BN>
int launch(const evo_gemm_params* p, cudaStream_t st, const long long* targets = nullptr, float4* part = nullptr) {
  using C_ = Cfg<CG, BN>;
  CUtensorMap tmA, tmB;
  int rc;
  int a_rows = BM;
  if (BN == BN_SMALL) { a_rows = 16; while (a_rows < BM && a_rows < p->M) a_rows *= 2; }
  if ((rc = make_tmap_2d_bf16(&tmA, p->A, (uint64_t)p->K, (uint64_t)p->M, (uint64_t)p->lda * 2, BK, (uint32_t)a_rows, true))) return rc;
  const bool b_tiled = BN == BN_SMALL && p->variant == 3;
  if (b_tiled) {
    uint64_t dims[4] = {(uint64_t)BK, (uint64_t)BN, (uint64_t)(p->K / BK), (uint64_t)(p->N / BN)};
    uint64_t str[3] = {(uint64_t)BK * 2, (uint64_t)BK * BN * 2, (uint64_t)(p->K / BK) * BK * BN * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)BN, 1, 1};
    if ((rc = make_tmap_nd_bf16(&tmB, p->W, 4, dims, str, box, true))) return rc;
  } else if ((rc = make_tmap_2d_bf16(&tmB, p->W, (uint64_t)p->K, (uint64_t)p->N, (uint64_t)p->K * 2, BK, C_::B_ROWS, true))) return rc;
  GemmArgs g;
  g.C = (bf16*)p->C; g.ldc = p->ldc; g.bias = (const bf16*)p->bias; g.resid = (const bf16*)p->residual; g.ldr = p->ldr;
  g.M = p->M; g.N = p->N; g.K = p->K;
  g.targets = targets; g.part = part;
  g.n_peers = p->n_c_peers; g.peer_period = (int)p->peer_period; g.peer_inner = (int)p->peer_inner; g.peer_row0 = p->peer_row0;
  for (int i = 0; i < 8; ++i) g.c_peer[i] = i < p->n_c_peers ? (bf16*)p->c_peers[i] : nullptr;
  g.rope_cos = (const bf16*)p->rope_cos; g.rope_sin = (const bf16*)p->rope_sin; g.rope_L = p->rope_L > 0 ? p->rope_L : 1; g.rope_cols = p->rope_cols;
  g.a_rows = a_rows;
  g.b_tiled = b_tiled;
  // small tile: two 64-column k-blocks per barrier round when K allows (halves the issue thread's serial rounds)
  g.ksub = (BN == BN_SMALL && p->K % (2 * BK) == 0 && a_rows <= 32) ? 2 : 1;
  g.n_stages = BN == BN_SMALL ? (a_rows <= 32 ? 8 / g.ksub : 4) : C_::STAGES;
  g.m_blocks = (int)((p->M + BM * CG - 1) / (BM * CG));
  g.n_blocks = (int)(p->N / BN);
  // Rasterisation: keep the smaller operand slab resident in L2 while the other one streams.
  //   grouped along M: GM row-blocks (GM * rows * K * 2 bytes of A) stay hot, W streams once per group;
  //   grouped along N: GN column-blocks of W stay hot, A streams once per group.
  // Pick the variant with the lower DRAM traffic estimate for a ~40 MB resident slab.
  {
    const double a_blk = (double)BM * CG * p->K * 2, w_blk = (double)BN * p->K * 2;
    const double a_tot = (double)p->M * p->K * 2, w_tot = (double)p->N * p->K * 2, budget = 40e6;
    // a short last group (16 column-blocks in groups of 7 leave 2) sweeps the other operand almost unshared: use one group
    // fewer when that stretches the slab by <= 20 %, else spread the blocks evenly (profiles/r02_gemm_raster_sweep_call25.txt:
    // K = 11008, groups 7+7+2 -> 8+8: DRAM reads 12.2 -> 10.2 GB, 2.4 % faster under ncu)
    static const char* env_b = getenv("EVO_B200_GEMM_REBALANCE");     // experiments only: 0 restores the plain budget rule
    auto balanced = [](int blocks, int gmax) {
      if (env_b && atoi(env_b) == 0) return gmax;
      const int groups = (blocks + gmax - 1) / gmax, last = blocks - (groups - 1) * gmax;
      if (groups == 1 || 2 * last >= gmax) return gmax;
      const int fewer = (blocks + groups - 2) / (groups - 1);
      return 5 * fewer <= 6 * gmax ? fewer : (blocks + groups - 1) / groups;
    };
    int gm = balanced(g.m_blocks, (int)std::max(1.0, std::min((double)g.m_blocks, budget / a_blk)));
    int gn = balanced(g.n_blocks, (int)std::max(1.0, std::min((double)g.n_blocks, budget / w_blk)));
    const double traffic_m = a_tot + w_tot * std::ceil((double)g.m_blocks / gm);
    const double traffic_n = w_tot + a_tot * std::ceil((double)g.n_blocks / gn);
    g.raster_n = traffic_n < traffic_m;
    g.group_m = g.raster_n ? gn : gm;
    static const char* env_h = getenv("EVO_B200_GEMM_L2_HINTS");   // experiments only: 0 (default) plain TMA loads, 1 / 2 see the producer
    g.l2_hints = (BN == BN_BIG && env_h) ? atoi(env_h) : 0;
    const char* env_s = getenv("EVO_B200_GEMM_SKEW");               // experiments only (read per launch)
    g.skew = (BN == BN_BIG && env_s) ? atoi(env_s) : 0;
    const char* env_g = getenv("EVO_B200_GEMM_GROUP");      // experiments only (read per launch: tools/gemm_raster_sweep.py changes them)
    const char* env_r = getenv("EVO_B200_GEMM_RASTER_N");
    if (env_r) g.raster_n = atoi(env_r);
    if (env_g) g.group_m = std::max(1, atoi(env_g));
  }
  static unsigned long long attr_done = 0;     // one bit per device (the template instance has its own copy)
  auto kern = gemm_tcgen05_kernel<CG, EPI, BN>;
  { int rc_ = ensure_dyn_smem(kern, C_::SMEM_BYTES, attr_done); if (rc_) return rc_; }
  int sms = device_sm_count();
  int n_tiles = g.m_blocks * g.n_blocks;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = C_::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  g.die_tab = nullptr; g.die_claim = nullptr; g.m_split = 0; g.die_pairs0 = g.die_pairs1 = 0; g.die_collide = 0;
  if (CG == 2) {
    int pairs = min(n_tiles, sms / 2);
    // Die-aware rasterisation (see the kernel's tile walk; EVO_B200_GEMM_DIE_RASTER=0 turns it off, 2 is a test hook): a full
    // grid of pairs, enough row-blocks for both dies, and a self-consistent SM -> die map of this device.  m_split balances the two
    // dies' wave counts.  In-step +2.0 / +2.7 % on the 8k workload (profiles/r02_gemm_die_raster_call28.txt).
    const char* env_d = getenv("EVO_B200_GEMM_DIE_RASTER");          // read per launch (tools/gemm_raster_sweep.py toggles it)
    const int die_mode = env_d ? atoi(env_d) : 1;
    if (BN == BN_BIG && die_mode != 0 && pairs == sms / 2 && g.m_blocks >= 32) {
      const DieMap* dm = die_map(st);
      if (dm && dm->pairs[0] + dm->pairs[1] == pairs) {
        const double prop = (double)g.m_blocks * dm->pairs[0] / pairs;
        long long best = -1; int best_ms = 0;
        for (int ms = std::max(1, (int)prop - 2); ms <= std::min(g.m_blocks - 1, (int)prop + 3); ++ms) {
          const long long w0 = ((long long)ms * g.n_blocks + dm->pairs[0] - 1) / dm->pairs[0];
          const long long w1 = ((long long)(g.m_blocks - ms) * g.n_blocks + dm->pairs[1] - 1) / dm->pairs[1];
          const long long w = std::max(w0, w1) * 1024 + (long long)(std::fabs(ms - prop) * 16);     // fewest waves, then closest to proportional
          if (best < 0 || w < best) { best = w; best_ms = ms; }
        }
        // claim words of this launch: the next 128-word line of a 256-line ring (a line is reused 256 die-aware launches later)
        unsigned* claim = die_next_claims(dm);
        EVO_CUDA(cudaMemsetAsync(claim, 0, DieMap::CLAIM_WORDS * sizeof(unsigned), st));
        g.die_tab = dm->tab; g.die_claim = claim; g.m_split = best_ms; g.die_pairs0 = dm->pairs[0]; g.die_pairs1 = dm->pairs[1];
        g.die_collide = die_mode == 2;
      }
    }
    cfg.gridDim = dim3(pairs * 2);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(min(n_tiles, sms * C_::CTAS_PER_SM));
  }
  EVO_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, g));
  return check_launch("evo_gemm");
}

template <int CG, int BN>
int dispatch_epi(const evo_gemm_params* p, cudaStream_t st) {
  switch (p->epilogue) {
    case EVO_EPI_NONE: return launch<CG, EVO_EPI_NONE, BN>(p, st);
    case EVO_EPI_BIAS: return launch<CG, EVO_EPI_BIAS, BN>(p, st);
    case EVO_EPI_BIAS_RESID: return launch<CG, EVO_EPI_BIAS_RESID, BN>(p, st);
    case EVO_EPI_RESID: return launch<CG, EVO_EPI_RESID, BN>(p, st);
    case EVO_EPI_BIAS_ROPE:
      if constexpr (BN == BN_BIG) return launch<CG, EVO_EPI_BIAS_ROPE, BN>(p, st);
      else { set_error("evo_gemm: the rotary epilogue is not available on the small-M tile (variant 2)"); return -1; }
    case EVO_EPI_GELU_GATE:
      if constexpr (BN == BN_BIG) return launch<CG, EVO_EPI_GELU_GATE, BN>(p, st);
      else { set_error("evo_gemm: the GELU-gate epilogue is not available on the small-M tile (variant 2)"); return -1; }
  }
  set_error("evo_gemm: unknown epilogue %d", p->epilogue);
  return -1;
}

}  // namespace

extern "C" int evo_gemm(const evo_gemm_params* p, void* stream) {
  EVO_REQUIRE(p->M >= 0 && p->N > 0 && p->K > 0, "evo_gemm: bad shape");
  EVO_REQUIRE(p->K % BK == 0, "evo_gemm: K (%lld) must be a multiple of %d", (long long)p->K, BK);
  EVO_REQUIRE(p->N % BN_BIG == 0, "evo_gemm: N (%lld) must be a multiple of %d (pack weights at load time)", (long long)p->N, BN_BIG);
  EVO_REQUIRE(p->lda % 8 == 0 && p->ldc % 8 == 0, "evo_gemm: lda/ldc must be multiples of 8 elements");
  EVO_REQUIRE(((uintptr_t)p->A % 16) == 0 && ((uintptr_t)p->W % 16) == 0 && ((uintptr_t)p->C % 16) == 0, "evo_gemm: pointers must be 16-byte aligned");
  if (p->epilogue == EVO_EPI_BIAS || p->epilogue == EVO_EPI_BIAS_RESID || p->epilogue == EVO_EPI_BIAS_ROPE) EVO_REQUIRE(p->bias != nullptr, "evo_gemm: bias epilogue without bias");
  if (p->epilogue == EVO_EPI_BIAS_ROPE)
    EVO_REQUIRE(p->rope_cos && p->rope_sin && p->rope_L > 0 && p->rope_cols % 128 == 0 && p->rope_cols <= p->N && ((uintptr_t)p->rope_cos % 16) == 0 && ((uintptr_t)p->rope_sin % 16) == 0,
                "evo_gemm: rotary epilogue needs 16-byte aligned cos/sin tables, rope_L > 0 and rope_cols a multiple of the head width (128)");
  if (p->epilogue == EVO_EPI_RESID || p->epilogue == EVO_EPI_BIAS_RESID)
    EVO_REQUIRE(p->residual != nullptr && p->ldr % 8 == 0 && ((uintptr_t)p->residual % 16) == 0, "evo_gemm: residual epilogue without a valid residual");
  if (p->n_c_peers != 0) {
    EVO_REQUIRE(p->n_c_peers > 0 && p->n_c_peers <= 8 && p->c_peers != nullptr, "evo_gemm: 1..8 peer destinations");
    EVO_REQUIRE(p->peer_inner > 0 && p->peer_inner % 128 == 0 && p->peer_period == p->peer_inner * p->n_c_peers && p->N % p->peer_period == 0,
                "evo_gemm: peer scatter needs peer_inner %% 128 == 0, peer_period = n_peers * peer_inner, N %% peer_period == 0");
    EVO_REQUIRE(p->epilogue == EVO_EPI_BIAS || p->epilogue == EVO_EPI_NONE || p->epilogue == EVO_EPI_BIAS_ROPE, "evo_gemm: peer scatter supports the none / bias / rotary epilogues");
    EVO_REQUIRE(p->variant == 0 || p->variant == 1, "evo_gemm: peer scatter needs the 256-column tiles");
  }
  if (p->M == 0) return 0;
  if (p->variant == 2 || p->variant == 3) return dispatch_epi<1, BN_SMALL>(p, (cudaStream_t)stream);
  if (p->variant == 1) return dispatch_epi<1, BN_BIG>(p, (cudaStream_t)stream);
  return dispatch_epi<2, BN_BIG>(p, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Fused scoring head: unembed GEMM + log-softmax + gather (+ entropy); the (M, V) logits never reach HBM.
// The GEMM's epilogue leaves one float4 of statistics per (row, 256-column half); the finish kernel folds the halves:
//   lse = M + log(sum_j s_j e^(m_j - M)),  logprob = x_target - lse,  entropy = lse - (sum_j w_j e^(m_j - M)) / S.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void score_finish_kernel(const float4* __restrict__ part, int n_blocks, long long M, const long long* __restrict__ targets, int V,
                                    float* __restrict__ logprobs, float* __restrict__ entropy) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float mx = -INFINITY;
  for (int j = 0; j < n_blocks; ++j) mx = fmaxf(mx, part[row * n_blocks + j].x);
  float S = 0.f, W = 0.f;
  for (int j = 0; j < n_blocks; ++j) {
    const float4 q = part[row * n_blocks + j];
    const float sc = expf(q.x - mx);
    S = fmaf(q.y, sc, S); W = fmaf(q.z, sc, W);
  }
  const float lse = mx + logf(S);
  if (logprobs) {
    const long long t = targets ? targets[row] : -1;
    logprobs[row] = (t < 0 || t >= V) ? 0.f : part[row * n_blocks + (int)(t / BN_BIG)].w - lse;
  }
  if (entropy) entropy[row] = lse - W / S;
}
}  // namespace

extern "C" size_t evo_unembed_score_workspace(int64_t M, int V) { return (size_t)M * (size_t)(V / BN_BIG) * sizeof(float4); }

extern "C" int evo_unembed_score(const evo_score_params* p, void* stream) {
  EVO_REQUIRE(p->M >= 0 && p->V > 0 && p->V % BN_BIG == 0, "evo_unembed_score: vocabulary (%d) must be a multiple of %d", p->V, BN_BIG);
  EVO_REQUIRE(p->K > 0 && p->K % BK == 0, "evo_unembed_score: K (%lld) must be a multiple of %d", (long long)p->K, BK);
  EVO_REQUIRE(((uintptr_t)p->x % 16) == 0 && ((uintptr_t)p->W % 16) == 0, "evo_unembed_score: pointers must be 16-byte aligned");
  EVO_REQUIRE(p->logprobs == nullptr || p->targets != nullptr, "evo_unembed_score: logprobs need targets");
  if (p->M == 0) return 0;
  const size_t need = evo_unembed_score_workspace(p->M, p->V);
  EVO_REQUIRE(p->workspace != nullptr && p->workspace_bytes >= need, "evo_unembed_score: workspace too small (%zu < %zu)", p->workspace_bytes, need);
  evo_gemm_params gp = {};
  gp.A = p->x; gp.lda = p->K; gp.W = p->W; gp.C = nullptr; gp.ldc = p->V; gp.M = p->M; gp.N = p->V; gp.K = p->K; gp.epilogue = EPI_LSE; gp.variant = 0;
  int rc = launch<2, EPI_LSE, BN_BIG>(&gp, (cudaStream_t)stream, (const long long*)p->targets, (float4*)p->workspace);
  if (rc) return rc;
  score_finish_kernel<<<(unsigned)((p->M + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)p->workspace, p->V / BN_BIG, p->M, (const long long*)p->targets, p->V,
                                                                                       p->logprobs, p->entropy);
  return check_launch("evo_unembed_score");
}
