// Causal attention, ping-pong variant (variant 2 of evo_attn_fwd_ws): two 128-row query tiles of the
// same (batch, head) per CTA, so the softmax of one tile runs on the CUDA cores / MUFU while the tensor
// pipe works on the other tile's QK^T and PV — the profile of the single-tile kernel (attention.cu)
// shows both pipes idle more than half of the time because a tile's softmax phases are serial.
//
//   warps 0-3   softmax + epilogue of tile A (thread r <-> query row r <-> TMEM lane r: no shuffles)
//   warps 4-7   softmax + epilogue of tile B
//   warp  8     TMA producer (Q_A, Q_B once; K / V tiles through 2-deep rings shared by both tiles)
//   warp  9     MMA issuer + TMEM owner
//   KV tiles are 64 keys.  TMEM (512 columns): per query tile two S buffers of 64 fp32 columns and one O of
//   128.  P (bf16) is written back INTO the S buffer it came from (32 columns, tcgen05.st) and consumed by
//   the PV product as the A operand straight from TMEM, so P never touches shared memory; V is read in place
//   as an MN-major B operand.  Because S is double-buffered per tile, QK(j+1) is issued BEFORE softmax(j)
//   finishes (the first cut aliased P onto a single S buffer and its softmax warps spent a third of their
//   time waiting for S); issue order: QK(0) QK(1) | PV(j) QK(j+2) for j = 0.. , tiles A and B interleaved.
#include "common.cuh"
#include "../../include/evo_b200.h"

using namespace evo;

namespace evo_attn_pp {

constexpr int HD = 128, BQ = 128, BKV = 64;
constexpr int TILE_BYTES = 128 * 128 * 2, HALF_BYTES = TILE_BYTES / 2;   // a Q tile: 2 d-halves of 128 rows x 64
constexpr int KV_TILE = BKV * HD * 2, KV_HALF = KV_TILE / 2;                   // a K or V tile: 2 d-halves of 64 keys x 64
constexpr int KV_STAGES = 4;
constexpr int THREADS = 320;
constexpr int SMEM = 2 * TILE_BYTES + 2 * KV_STAGES * KV_TILE + 256;
static_assert(SMEM <= 232448, "shared memory budget");
constexpr uint32_t TM_S = 0, TM_O = 256;        // S: + 128 * tile + 64 * buffer;  O: + 128 * tile

struct Args {
  bf16* out;
  int B, H;
  long long Lq, Lk, q_pos0;
  float scale_log2;
  int n_qblk;
};

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 2^x for two values on the FMA pipe (the MUFU pipe, 16 ex2/clk/SM, is the co-bottleneck of this kernel):
// round-to-nearest split x = n + f via the 1.5*2^23 trick, degree-3 minimax of 2^f on [-0.5, 0.5]
// (max relative error 7.7e-5, far below the bf16 rounding P gets anyway), exponent patched in with integer adds.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, nmagic);
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(0.05508868f, 0.05508868f), make_float2(0.24260405f, 0.24260405f));
  p = __ffma2_rn(p, f, make_float2(0.69327624f, 0.69327624f));
  p = __ffma2_rn(p, f, make_float2(0.99992894f, 0.99992894f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
attn_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                               // [2 tiles]
  uint8_t* sK = sQ + 2 * TILE_BYTES;                // [KV_STAGES]
  uint8_t* sV = sK + KV_STAGES * KV_TILE;           // [KV_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KV_STAGES * KV_TILE);
  uint64_t* q_full = bars;                          // 2
  uint64_t* k_full = q_full + 2;                    // KV_STAGES
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;           // [tile][buffer]
  uint64_t* p_full = s_full + 4;                    // [tile][buffer]
  uint64_t* o_full = p_full + 4;                    // 2
  uint64_t* pv_done = o_full + 2;                   // 2: completion of every PV of a tile (phase index = kv tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = (a.n_qblk + 1) / 2 - 1 - (int)blockIdx.x;       // longest pairs first
  const int h = blockIdx.y, b = blockIdx.z;
  long long q0[2];
  int n_kv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    q0[t] = (long long)(2 * pair + t) * BQ;
    if (q0[t] < a.Lq) {
      const long long last_key = min(a.Lk - 1, a.q_pos0 + q0[t] + BQ - 1);
      n_kv[t] = (int)(last_key / BKV) + 1;
    } else n_kv[t] = 0;
  }
  const int n_kv_max = max(n_kv[0], n_kv[1]);

  if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&o_full[i], 1); mbar_init(&pv_done[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 9) { __syncwarp(); tmem_alloc<1>(tmem_slot, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int t = 0; t < 2; ++t) {
        if (n_kv[t] == 0) continue;
        mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
        tma_load_4d(sQ + t * TILE_BYTES, &tmQ, &q_full[t], 0, h, (int)q0[t], b);
        tma_load_4d(sQ + t * TILE_BYTES + HALF_BYTES, &tmQ, &q_full[t], 64, h, (int)q0[t], b);
      }
      for (int j = 0; j < n_kv_max; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (uint32_t)(j / KV_STAGES) & 1;
        const int key0 = j * BKV;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], KV_TILE);
        tma_load_4d(sK + st * KV_TILE, &tmK, &k_full[st], 0, h, key0, b);
        tma_load_4d(sK + st * KV_TILE + KV_HALF, &tmK, &k_full[st], 64, h, key0, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], KV_TILE);
        tma_load_4d(sV + st * KV_TILE, &tmV, &v_full[st], 0, h, key0, b);
        tma_load_4d(sV + st * KV_TILE + KV_HALF, &tmV, &v_full[st], 64, h, key0, b);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD) | (1u << 16);     // B (= V) is MN-major
      auto last_user = [&](int t, int j) { return t == 1 || j >= n_kv[1]; };  // tile B is the later tile: n_kv[1] >= n_kv[0] when valid
      auto issue_qk = [&](int t, int j) {
        const int st = j % KV_STAGES;
        mbar_wait(&k_full[st], (uint32_t)(j / KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ + t * TILE_BYTES), ka = smem_u32(sK + st * KV_TILE);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_ss<1>(tmem_base + TM_S + t * 128 + (j & 1) * BKV, umma_desc_k_sw128(qa + (kk >> 2) * HALF_BYTES + (kk & 3) * 32),
                     umma_desc_k_sw128(ka + (kk >> 2) * KV_HALF + (kk & 3) * 32), idesc_qk, kk != 0);
        if (last_user(t, j)) umma_commit(&k_empty[st]);
        umma_commit(&s_full[t * 2 + (j & 1)]);
      };
      for (int t = 0; t < 2; ++t) if (n_kv[t] > 0) mbar_wait(&q_full[t], 0);
      for (int j = 0; j < 2; ++j)
        for (int t = 0; t < 2; ++t) if (j < n_kv[t]) issue_qk(t, j);
      for (int j = 0; j < n_kv_max; ++j) {
        const int st = j % KV_STAGES;
        for (int t = 0; t < 2; ++t) {
          if (j >= n_kv[t]) continue;
          mbar_wait(&v_full[st], (uint32_t)(j / KV_STAGES) & 1);
          mbar_wait(&p_full[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1);
          tc_fence_after();
          const uint32_t va = smem_u32(sV + st * KV_TILE);
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk)
            umma_ts(tmem_base + TM_O + t * HD, tmem_base + TM_S + t * 128 + (j & 1) * BKV + kk * 8,
                    desc_mn_sw128(va + kk * 16 * 128, KV_HALF, 1024), idesc_pv, (j | kk) != 0);
          if (last_user(t, j)) umma_commit(&v_empty[st]);
          umma_commit(&pv_done[t]);
          if (j == n_kv[t] - 1) umma_commit(&o_full[t]);
          if (j + 2 < n_kv[t]) issue_qk(t, j + 2);     // reuses the S buffer whose P the PV above has just been queued behind
        }
      }
    }
  } else {
    // ------------------------------------------------ softmax + epilogue: warps 0-3 tile A, 4-7 tile B
    const int t = warp >> 2, q = warp & 3;
    if (n_kv[t] > 0) {
      const int r = q * 32 + lane;
      const long long pos = a.q_pos0 + q0[t] + r;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const uint32_t o_addr = lane_addr + TM_O + t * HD;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv[t]; ++j) {
        const int sb = j & 1;
        const uint32_t s_addr = lane_addr + TM_S + t * 128 + sb * BKV;
        mbar_wait(&s_full[t * 2 + sb], (uint32_t)(j >> 1) & 1);
        tc_fence_after();
        const long long key0 = (long long)j * BKV;
        // ---- one TMEM read of this thread's 64 scores
        float sc[BKV];
        {
          uint32_t t0[32], t1[32];
          tmem_ld_32x32(s_addr, t0); tmem_ld_32x32(s_addr + 32, t1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { sc[i] = __uint_as_float(t0[i]); sc[32 + i] = __uint_as_float(t1[i]); }
        }
        if (key0 + BKV - 1 > pos) {
#pragma unroll
          for (int i = 0; i < BKV; ++i) if (key0 + i > pos) sc[i] = -INFINITY;
        }
        float mxs[8];
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) mxs[q8] = sc[q8];
#pragma unroll
        for (int i = 8; i < BKV; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], sc[i]);
        float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])), fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7]))) * a.scale_log2;
        float alpha = 1.f;
        bool grow = mx > m_ref + 8.f;
        if (j == 0) { m_ref = (mx == -INFINITY) ? 0.f : mx; grow = false; }
        else if (grow) { alpha = ex2(m_ref - mx); m_ref = mx; l *= alpha; }
        // P = exp2(s*scale - m_ref) as bf16, written over the first 32 columns of this S buffer
        float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p0 = ex2(fmaf(sc[2 * i], a.scale_log2, -m_ref)), p1 = ex2(fmaf(sc[2 * i + 1], a.scale_log2, -m_ref));
          if (i & 1) { ls2 += p0; ls3 += p1; } else { ls0 += p0; ls1 += p1; }
          w[i] = pack_bf16(p0, p1);
        }
        l += (ls0 + ls2) + (ls1 + ls3);
        tmem_st_32x32(s_addr, w);
        // O rescale (rare): PV(j-1) must have completed; with QK issued two tiles ahead s_full(j) only covers PV(j-2),
        // so the issuer commits pv_done[t] behind every PV (phase index == j; PV(j) cannot be queued before our arrive below)
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          mbar_wait(&pv_done[t], (uint32_t)(j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < HD; c += 32) {
            uint32_t tt[32];
            tmem_ld_32x32(o_addr + c, tt);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) tt[i] = __float_as_uint(__uint_as_float(tt[i]) * alpha);
            tmem_st_32x32(o_addr + c, tt);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t * 2 + sb]);
      }
      // ---- epilogue
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const float inv_l = 1.f / l;
      bf16* orow = a.out + (((long long)b * a.Lq + q0[t] + r) * a.H + h) * HD;
#pragma unroll 1
      for (int c = 0; c < HD; c += 32) {
        uint32_t tt[32];
        tmem_ld_32x32(o_addr + c, tt);
        tmem_ld_wait();
        if (q0[t] + r < a.Lq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(tt[8 * i + 0]) * inv_l, __uint_as_float(tt[8 * i + 1]) * inv_l);
            o.y = pack_bf16(__uint_as_float(tt[8 * i + 2]) * inv_l, __uint_as_float(tt[8 * i + 3]) * inv_l);
            o.z = pack_bf16(__uint_as_float(tt[8 * i + 4]) * inv_l, __uint_as_float(tt[8 * i + 5]) * inv_l);
            o.w = pack_bf16(__uint_as_float(tt[8 * i + 6]) * inv_l, __uint_as_float(tt[8 * i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + 8 * i) = o;
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace evo_attn_pp

// called from evo_attn_fwd_ws (attention.cu) for variant 2
int evo_attn_pp_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const evo_attn_params* p, cudaStream_t st) {
  using namespace evo_attn_pp;
  Args a;
  a.out = (bf16*)p->out; a.B = p->B; a.H = p->H; a.Lq = p->Lq; a.Lk = p->Lk; a.q_pos0 = p->q_pos0;
  a.scale_log2 = p->softmax_scale * 1.4426950408889634f;
  a.n_qblk = (int)((p->Lq + BQ - 1) / BQ);
  static bool done = false;
  if (!done) { EVO_CUDA(cudaFuncSetAttribute(attn_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); done = true; }
  dim3 grid((unsigned)((a.n_qblk + 1) / 2), p->H, p->B);
  attn_pp_kernel<<<grid, THREADS, SMEM, st>>>(tmQ, tmK, tmV, a);
  return check_launch("evo_attn_fwd(pp)");
}
