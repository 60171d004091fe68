// Causal attention, ping-pong variant (variant 2 of evo_attn_fwd_ws): two 128-row query tiles of the
// same (batch, head) per CTA, so the softmax of one tile runs on the CUDA cores / MUFU while the tensor
// pipe works on the other tile's QK^T and PV — the profile of the single-tile kernel (attention.cu)
// shows both pipes idle more than half of the time because a tile's softmax phases are serial.
//
//   warps 0-3   softmax + epilogue of tile A (thread r <-> query row r <-> TMEM lane r: no shuffles)
//   warps 4-7   softmax + epilogue of tile B
//   warp  8     TMA producer (Q_A, Q_B once; K / V tiles through 2-deep rings shared by both tiles)
//   warp  9     MMA issuer + TMEM owner
//   TMEM (512 columns): S_A | S_B | O_A | O_B, 128 fp32 columns each.  P (bf16) is written back INTO the
//   S columns it came from (64 columns, tcgen05.st) and consumed by the PV product as the A operand
//   straight from TMEM (tcgen05.mma with a TMEM A operand), so P never touches shared memory; V is read in
//   place as an MN-major B operand.
//   Issue order per KV tile j:  PV_A(j) QK_A(j+1) PV_B(j) QK_B(j+1): while softmax A(j+1) runs, the tensor
//   pipe has PV_B(j) and QK_B(j+1) to do, and vice versa.
//   Measured on B200 (profiles/): 0.89-1.07 PFLOP/s; tensor pipe 50 % active, softmax warps wait for S a third of
//   the time (P aliases S, so QK(j+1) queues behind PV(j)).  Tried and rejected: a second TMEM pass over S instead
//   of holding the row in registers (0.50-0.57 PFLOP/s: tcgen05.ld bandwidth), a degree-3 FMA-pipe exp2 for half of
//   the scores (0.82-0.96: issue slots, not MUFU, bind), 64-key tiles with double-buffered S (0.65-0.76).
#include "common.cuh"
#include "../../include/evo_b200.h"

using namespace evo;

namespace evo_attn_pp {

constexpr int HD = 128, BQ = 128, BKV = 128;
constexpr int TILE_BYTES = 128 * 128 * 2, HALF_BYTES = TILE_BYTES / 2;
constexpr int KV_STAGES = 2;
constexpr int THREADS = 320;
constexpr int SMEM = TILE_BYTES * (2 + 2 * KV_STAGES) + 256;
static_assert(SMEM <= 232448, "shared memory budget");
constexpr uint32_t TM_S = 0, TM_O = 256;        // + 128 * tile

struct Args {
  bf16* out;
  bf16* out_peer[8]; int n_out_peers; long long out_rows_per_peer, out_row_stride, out_col0;   // peer-scattered output (evo_attn_params)
  int B, H;
  long long Lq, Lk, q_pos0;
  float scale_log2;
  int n_qblk;
};

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

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

// 2^x for a pair of scores on the FMA pipe (variant 3): round-to-nearest range reduction with the 1.5*2^23 trick, a cubic
// in f = x - round(x) in [-0.5, 0.5] (near-minimax for the relative error: 7.5e-5, a fiftieth of the bf16 rounding P gets next),
// and the exponent added as an integer.  x is clamped at -126 (masked scores are -inf: they become 2^-126, which rounds to 0
// against any row sum).  7 packed fma-pipe instructions + 4 ALU per PAIR, against 2 MUFU.EX2 at 8 issue cycles each.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 xi = __fadd2_rn(t, nmagic);
  const float2 f = __ffma2_rn(xi, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(make_float2(0.0551716648f, 0.0551716648f), f, make_float2(0.2426111251f, 0.2426111251f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677f, 0.6932609677f));
  p = __ffma2_rn(p, f, make_float2(0.9999280572f, 0.9999280572f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)), __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// NPOLY: of every 8 score pairs, how many take exp2_poly2 instead of MUFU.EX2 (0 = variant 2; 3 = variant 3: MUFU and the FMA
// pipe finish a tile's exponentials at about the same time)
template <int NPOLY>
__global__ void __launch_bounds__(THREADS, 1)
attn_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                               // [2 tiles]
  uint8_t* sK = sQ + 2 * TILE_BYTES;                // [KV_STAGES]
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;        // [KV_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KV_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;                          // 2
  uint64_t* k_full = q_full + 2;                    // KV_STAGES
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;           // 2 (per query tile)
  uint64_t* p_full = s_full + 2;                    // 2
  uint64_t* o_full = p_full + 2;                    // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

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
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1); }
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
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_4d(sK + st * TILE_BYTES, &tmK, &k_full[st], 0, h, key0, b);
        tma_load_4d(sK + st * TILE_BYTES + HALF_BYTES, &tmK, &k_full[st], 64, h, key0, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_4d(sV + st * TILE_BYTES, &tmV, &v_full[st], 0, h, key0, b);
        tma_load_4d(sV + st * TILE_BYTES + HALF_BYTES, &tmV, &v_full[st], 64, h, key0, b);
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
        const uint32_t qa = smem_u32(sQ + t * TILE_BYTES), ka = smem_u32(sK + st * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_ss<1>(tmem_base + TM_S + t * BKV, umma_desc_k_sw128(qa + off), umma_desc_k_sw128(ka + off), idesc_qk, kk != 0);
        }
        if (last_user(t, j)) umma_commit(&k_empty[st]);
        umma_commit(&s_full[t]);
      };
      for (int t = 0; t < 2; ++t) if (n_kv[t] > 0) { mbar_wait(&q_full[t], 0); issue_qk(t, 0); }
      for (int j = 0; j < n_kv_max; ++j) {
        const int st = j % KV_STAGES;
        for (int t = 0; t < 2; ++t) {
          if (j >= n_kv[t]) continue;
          mbar_wait(&v_full[st], (uint32_t)(j / KV_STAGES) & 1);
          mbar_wait(&p_full[t], (uint32_t)j & 1);
          tc_fence_after();
          const uint32_t va = smem_u32(sV + st * TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk)
            umma_ts(tmem_base + TM_O + t * HD, tmem_base + TM_S + t * BKV + kk * 8, desc_mn_sw128(va + kk * 16 * 128, HALF_BYTES, 1024), idesc_pv, (j | kk) != 0);
          if (last_user(t, j)) umma_commit(&v_empty[st]);
          if (j == n_kv[t] - 1) umma_commit(&o_full[t]);
          if (j + 1 < n_kv[t]) issue_qk(t, j + 1);
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
      const uint32_t s_addr = lane_addr + TM_S + t * BKV, o_addr = lane_addr + TM_O + t * HD;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv[t]; ++j) {
        mbar_wait(&s_full[t], (uint32_t)j & 1);
        tc_fence_after();
        const long long key0 = (long long)j * BKV;
        const bool need_mask = key0 + BKV - 1 > pos;
        // ---- one TMEM read of the row (tcgen05.ld moves ~64 B/clk: a second pass over S costs as much as the MMAs)
        float sc[BKV];
        {
          uint32_t t0[32], t1[32], t2[32], t3[32];
          tmem_ld_32x32(s_addr, t0); tmem_ld_32x32(s_addr + 32, t1); tmem_ld_32x32(s_addr + 64, t2); tmem_ld_32x32(s_addr + 96, t3);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            sc[i] = __uint_as_float(t0[i]); sc[32 + i] = __uint_as_float(t1[i]);
            sc[64 + i] = __uint_as_float(t2[i]); sc[96 + i] = __uint_as_float(t3[i]);
          }
        }
        if (need_mask) {
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
        // O rescale: s_full(j) already implies PV(j-1) has completed (QK(j) was issued behind it)
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
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
        // ---- P = exp2(s*scale - m_ref) as bf16, written over the S columns (all of S is in registers by now)
        float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
        if constexpr (NPOLY == 0) {
#pragma unroll
          for (int c = 0; c < BKV; c += 32) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2(fmaf(sc[c + 2 * i], a.scale_log2, -m_ref)), p1 = ex2(fmaf(sc[c + 2 * i + 1], a.scale_log2, -m_ref));
              if (i & 1) { ls2 += p0; ls3 += p1; } else { ls0 += p0; ls1 += p1; }
              w[i] = pack_bf16(p0, p1);
            }
            tmem_st_32x32_x16(s_addr + c / 2, w);
          }
        } else {
          // packed arithmetic (one FFMA2 scales two scores, one FADD2 sums two probabilities) and NPOLY of every 8 pairs
          // exponentiated on the FMA pipe, interleaved with the MUFU pairs so that both pipes stay busy inside one warp
          const float2 scale2 = make_float2(a.scale_log2, a.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
          float2 la = make_float2(0.f, 0.f), lb = la;
#pragma unroll
          for (int c = 0; c < BKV; c += 32) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 x = __ffma2_rn(make_float2(sc[c + 2 * i], sc[c + 2 * i + 1]), scale2, nm2);
              float2 p;
              if ((i & 7) * NPOLY % 8 < NPOLY) p = exp2_poly2(x);      // NPOLY pairs of 8, spread evenly
              else p = make_float2(ex2(x.x), ex2(x.y));
              if (i & 1) lb = __fadd2_rn(lb, p); else la = __fadd2_rn(la, p);
              w[i] = pack_bf16(p.x, p.y);
            }
            tmem_st_32x32_x16(s_addr + c / 2, w);
          }
          ls0 = la.x + lb.x; ls1 = la.y + lb.y;
        }
        const float ls0_ = ls0 + ls2, ls1_ = ls1 + ls3;
        ls0 = ls0_; ls1 = ls1_;
        l += ls0 + ls1;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
      }
      // ---- epilogue
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const float inv_l = 1.f / l;
      bf16* orow = a.out + (((long long)b * a.Lq + q0[t] + r) * a.H + h) * HD;
      if (a.n_out_peers > 0) {
        const long long tok = q0[t] + r, pr = tok / a.out_rows_per_peer;
        orow = a.out_peer[pr < a.n_out_peers ? pr : 0] + (tok - pr * a.out_rows_per_peer) * a.out_row_stride + a.out_col0 + (long long)h * HD;
      }
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
int evo_attn_pp_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const evo_attn_params* p, cudaStream_t st, int npoly) {
  using namespace evo_attn_pp;
  Args a;
  a.out = (bf16*)p->out; a.B = p->B; a.H = p->H; a.Lq = p->Lq; a.Lk = p->Lk; a.q_pos0 = p->q_pos0;
  a.scale_log2 = p->softmax_scale * 1.4426950408889634f;
  a.n_qblk = (int)((p->Lq + BQ - 1) / BQ);
  a.n_out_peers = p->n_out_peers; a.out_rows_per_peer = p->out_rows_per_peer; a.out_row_stride = p->out_row_stride; a.out_col0 = p->out_col0;
  for (int i = 0; i < 8; ++i) a.out_peer[i] = i < p->n_out_peers ? (bf16*)p->out_peers[i] : nullptr;
  if (p->n_out_peers != 0) {
    EVO_REQUIRE(p->n_out_peers > 0 && p->n_out_peers <= 8 && p->out_peers != nullptr && p->B == 1 && p->out_rows_per_peer > 0 &&
                p->out_rows_per_peer * p->n_out_peers >= p->Lq && p->out_row_stride % 8 == 0 && p->out_col0 % 8 == 0,
                "evo_attn_fwd: peer-scattered output needs B == 1, 1..8 peers covering Lq rows, 16-byte aligned strides");
  }
  static unsigned long long done0 = 0, done3 = 0;
  dim3 grid((unsigned)((a.n_qblk + 1) / 2), p->H, p->B);
  if (npoly == 0) {
    { int rc_ = ensure_dyn_smem(attn_pp_kernel<0>, SMEM, done0); if (rc_) return rc_; }
    attn_pp_kernel<0><<<grid, THREADS, SMEM, st>>>(tmQ, tmK, tmV, a);
  } else {
    { int rc_ = ensure_dyn_smem(attn_pp_kernel<3>, SMEM, done3); if (rc_) return rc_; }
    attn_pp_kernel<3><<<grid, THREADS, SMEM, st>>>(tmQ, tmK, tmV, a);
  }
  return check_launch("evo_attn_fwd(pp)");
}
