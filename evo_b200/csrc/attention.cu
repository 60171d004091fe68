// Causal rotary attention core for sm_100a (head_dim 128): softmax(Q K^T / sqrt(d)) V.
// Replaces flash_attn_qkvpacked_func (flash_attn/modules/mha.py:122; FA2 = mma.sync HMMA
// kernels recompiled for sm_100) with a tcgen05/TMEM kernel:
//
//   one CTA = 128 query rows of one (batch, head); 10 warps:
//     warps 0-7  softmax + epilogue: thread (q, lane) of column-half hf owns query row
//                r = 32q + lane (== TMEM lane r) and 64 of the tile's 128 key columns, so a row
//                needs no shuffles, only one smem exchange of the half-row max per tile; two
//                warps per SM sub-partition hide the MUFU / TMEM / smem latencies of each other
//     warp 8     TMA producer: Q once, K and V tiles (128 keys) through 2-deep mbarrier rings
//     warp 9     MMA issuer: S = Q K^T (SS, 128x128x128) into one of two TMEM S buffers and
//                O += P V (SS, P staged by the softmax warps in 128B-swizzled smem)
//   S is double-buffered so Q K_{j+1}^T runs on the tensor pipe while the CUDA cores do
//   softmax(j); O stays in TMEM across KV tiles and is rescaled lazily (only when the running
//   max grows by more than 2^8), the final 1/l normalisation happens in the epilogue.
//
// Arithmetic matches FlashAttention's: fp32 scores and statistics, P rounded to bf16 for the
// PV product, fp32 accumulation, single bf16 rounding of the output.
#include "common.cuh"
#include "../../include/evo_b200.h"

using namespace evo;

namespace {

constexpr int HD = 128;          // head dim
constexpr int BQ = 128;          // query rows per CTA
constexpr int BKV = 128;         // keys per tile
constexpr int TILE_BYTES = 128 * 128 * 2;          // any 128x128 bf16 tile = 2 x (128 rows x 64) swizzled halves
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int ATT_THREADS = 320;
constexpr int SM_WARPS = 8;             // softmax warps; producer = warp 8, MMA = warp 9
constexpr int KV_STAGES = 2;
// 7 tiles + barriers + max-exchange buffer = 226.25 KB of the 227 KB limit: the dynamic smem base is
// declared 1024-aligned instead of padding for a manual round-up
constexpr int ATT_SMEM = TILE_BYTES * (1 + 2 * KV_STAGES + 2) + 256 + 2 * 2 * 128 * 4;
static_assert(ATT_SMEM <= 232448, "exceeds the 227 KB dynamic shared memory limit");   // Q + K ring + V ring + 2 P buffers
constexpr uint32_t TM_S0 = 0, TM_O = 256;

struct AttArgs {
  bf16* out;
  int B, H;
  long long Lq, Lk, q_pos0;
  float scale_log2;              // softmax_scale * log2(e)
};

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MN-major (N contiguous) B operand: tile staged as [2 d-halves][128 keys][64 d] with 128B swizzle
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// V_MN = false: V comes pre-transposed (B, H, 128, Lk_pad) and is a K-major B operand
// V_MN = true : V is read in place (keys x d) as an MN-major B operand
template <bool V_MN>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;                                  // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();               // fail loudly, never compute on a misaligned tile
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;
  uint8_t* sP = sV + KV_STAGES * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // KV_STAGES
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;  // 2
  uint64_t* s_empty = s_full + 2;          // 2
  uint64_t* p_full = s_empty + 2;          // 2
  uint64_t* pv_done = p_full + 2;          // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xch = reinterpret_cast<float*>(bars + 32);          // [2 parities][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qblk = (int)((a.Lq + BQ - 1) / BQ);
  const int qblk = n_qblk - 1 - (int)blockIdx.x;      // longest (latest) blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  const long long q0 = (long long)qblk * BQ;
  // keys visible to this block: j <= q_pos0 + q0 + 127, j < Lk
  const long long last_key = min(a.Lk - 1, a.q_pos0 + q0 + BQ - 1);
  const int n_kv = (int)(last_key / BKV) + 1;

  if (warp == SM_WARPS && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == SM_WARPS + 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 32 * SM_WARPS); mbar_init(&p_full[i], 32 * SM_WARPS); mbar_init(&pv_done[i], 1); }
    fence_barrier_init();
  }
  if (warp == SM_WARPS + 1) { __syncwarp(); tmem_alloc<1>(tmem_slot, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == SM_WARPS) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, (int)q0, b);
      tma_load_4d(sQ + HALF_BYTES, &tmQ, q_full, 64, h, (int)q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (uint32_t)(j / KV_STAGES) & 1;
        const int key0 = j * BKV;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_4d(sK + st * TILE_BYTES, &tmK, &k_full[st], 0, h, key0, b);
        tma_load_4d(sK + st * TILE_BYTES + HALF_BYTES, &tmK, &k_full[st], 64, h, key0, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        if constexpr (V_MN) {   // (d, H, key, B): two d-halves of 128 keys x 64 d
          tma_load_4d(sV + st * TILE_BYTES, &tmV, &v_full[st], 0, h, key0, b);
          tma_load_4d(sV + st * TILE_BYTES + HALF_BYTES, &tmV, &v_full[st], 64, h, key0, b);
        } else {                // (key, d, H, B): two key-halves of 128 d x 64 keys
          tma_load_4d(sV + st * TILE_BYTES, &tmV, &v_full[st], key0, 0, h, b);
          tma_load_4d(sV + st * TILE_BYTES + HALF_BYTES, &tmV, &v_full[st], key0 + 64, 0, h, b);
        }
      }
    }
  } else if (warp == SM_WARPS + 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD) | (V_MN ? (1u << 16) : 0u);
      auto issue_qk = [&](int j) {
        const int st = j % KV_STAGES, sb = j & 1;
        mbar_wait(&k_full[st], (uint32_t)(j / KV_STAGES) & 1);
        mbar_wait(&s_empty[sb], ((uint32_t)(j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + st * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_ss<1>(tmem_base + TM_S0 + sb * BKV, umma_desc_k_sw128(qa + off), umma_desc_k_sw128(ka + off), idesc_qk, kk != 0);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[sb]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_qk(j + 1);
        const int st = j % KV_STAGES, pb = j & 1;
        mbar_wait(&v_full[st], (uint32_t)(j / KV_STAGES) & 1);
        mbar_wait(&p_full[pb], (uint32_t)(j >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP + pb * TILE_BYTES), va = smem_u32(sV + st * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint32_t aoff = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          uint64_t bd;
          if constexpr (V_MN) bd = umma_desc_mn_sw128(va + kk * 16 * 128, HALF_BYTES, 1024);
          else                bd = umma_desc_k_sw128(va + aoff);
          umma_ss<1>(tmem_base + TM_O, umma_desc_k_sw128(pa + aoff), bd, idesc_pv, (j | kk) != 0);
        }
        umma_commit(&v_empty[st]);
        umma_commit(&pv_done[pb]);
      }
    }
  } else {
    // ------------------------------------------------ softmax + epilogue (warps 0-7)
    const int q = warp & 3, hf = warp >> 2;                 // TMEM lane quarter, key-column half
    const int r = q * 32 + lane;                            // query row in the block == TMEM lane
    const long long pos = a.q_pos0 + q0 + r;                // absolute position of this query
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int HC = BKV / 2;                             // 64 columns per thread
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      float s[HC];
      {
        uint32_t t0[32], t1[32];
        const uint32_t sa = lane_addr + TM_S0 + sb * BKV + hf * HC;
        tmem_ld_32x32(sa, t0); tmem_ld_32x32(sa + 32, t1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(t0[i]); s[32 + i] = __uint_as_float(t1[i]); }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);
      // causal / length mask on tiles that reach past this row's position
      const long long key0 = (long long)j * BKV + hf * HC;
      if (key0 + HC - 1 > pos) {
#pragma unroll
        for (int i = 0; i < HC; ++i) if (key0 + i > pos) s[i] = -INFINITY;
      }
      float mxs[8];
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8) mxs[q8] = s[q8];
#pragma unroll
      for (int i = 8; i < HC; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], s[i]);
      float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])), fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      // exchange the half-row max with the partner thread (same row, other column half)
      float* xb = xch + (j & 1) * 256;
      xb[hf * 128 + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, xb[(hf ^ 1) * 128 + r]) * a.scale_log2;
      // lazy rescale: keep the old reference unless the max grew by more than 2^8
      float alpha = 1.f;
      bool grow = mx > m_ref + 8.f;
      if (j == 0) { m_ref = (mx == -INFINITY) ? 0.f : mx; grow = false; }
      else if (grow) { alpha = ex2(m_ref - mx); m_ref = mx; l *= alpha; }
      // P = exp2(s*scale - m_ref), written as bf16 into this half of the swizzled A-operand tile
      uint8_t* prow = sP + sb * TILE_BYTES + hf * HALF_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 8; ++c) {                           // 16-byte chunks: 8 keys each
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float p0 = ex2(fmaf(s[c * 8 + 2 * i], a.scale_log2, -m_ref));
          float p1 = ex2(fmaf(s[c * 8 + 2 * i + 1], a.scale_log2, -m_ref));
          ls[i] += p0 + p1;
          w[i] = pack_bf16(p0, p1);
        }
        *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      // O rescale of this thread's 64 output columns (warp-collective because tcgen05.ld/st are)
      if (j > 0) {
        const bool any = __any_sync(0xffffffffu, grow);
        mbar_wait(&pv_done[(j - 1) & 1], (uint32_t)((j - 1) >> 1) & 1);
        if (any) {
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < HC; c += 32) {
            uint32_t t[32];
            tmem_ld_32x32(lane_addr + TM_O + hf * HC + c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st_32x32(lane_addr + TM_O + hf * HC + c, t);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      }
      fence_proxy_async_smem();      // P (generic-proxy stores) -> visible to the tensor core's async proxy
      mbar_arrive(&p_full[sb]);
    }
    // epilogue: O / l -> bf16 -> out[b, q0+r, h*128 + hf*64 : +64]; l = sum of both halves
    float* xb = xch + (n_kv & 1) * 256;
    xb[hf * 128 + r] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xb[(hf ^ 1) * 128 + r];
    mbar_wait(&pv_done[(n_kv - 1) & 1], (uint32_t)((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    bf16* orow = a.out + (((long long)b * a.Lq + q0 + r) * a.H + h) * HD + hf * HC;
#pragma unroll 1
    for (int c = 0; c < HC; c += 32) {
      uint32_t t[32];
      tmem_ld_32x32(lane_addr + TM_O + hf * HC + c, t);
      tmem_ld_wait();
      if (q0 + r < a.Lq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(t[8 * i + 0]) * inv_l, __uint_as_float(t[8 * i + 1]) * inv_l);
          o.y = pack_bf16(__uint_as_float(t[8 * i + 2]) * inv_l, __uint_as_float(t[8 * i + 3]) * inv_l);
          o.z = pack_bf16(__uint_as_float(t[8 * i + 4]) * inv_l, __uint_as_float(t[8 * i + 5]) * inv_l);
          o.w = pack_bf16(__uint_as_float(t[8 * i + 6]) * inv_l, __uint_as_float(t[8 * i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c + 8 * i) = o;
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS + 1) tmem_dealloc<1>(tmem_base, 512);
}

// (B, L, H, 128) strided -> (B, H, 128, Lpad): 64x64 smem tile transpose
__global__ void transpose_v_kernel(const bf16* __restrict__ v, bf16* __restrict__ vt, long long L, long long Lpad, int H,
                                   long long tok_stride, long long batch_stride) {
  __shared__ bf16 tile[64][66];
  const int h = blockIdx.y, b = blockIdx.z;
  const long long l0 = (long long)blockIdx.x * 64;
  for (int dh = 0; dh < HD; dh += 64) {
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
      int lr = i / 64, d = i % 64;
      long long l = l0 + lr;
      tile[lr][d] = l < L ? v[b * batch_stride + l * tok_stride + (long long)h * HD + dh + d] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
      int d = i / 64, lr = i % 64;
      long long l = l0 + lr;
      if (l < Lpad) vt[(((long long)b * H + h) * HD + dh + d) * Lpad + l] = tile[lr][d];
    }
    __syncthreads();
  }
}

typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3], const uint32_t box[4]) {
  static encode_fn_t fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return -1;
    }
    fn = (encode_fn_t)p;
  }
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, s, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EVO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed (%d): dims=%llu,%llu,%llu,%llu strides=%llu,%llu,%llu",
              (int)r, (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], (unsigned long long)d[3],
              (unsigned long long)s[0], (unsigned long long)s[1], (unsigned long long)s[2]);
  return 0;
}

}  // namespace

int evo_attn_pp_launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const evo_attn_params* p, cudaStream_t st, int npoly);   // attention_pp.cu

extern "C" size_t evo_attn_fwd_workspace(const evo_attn_params* p, int variant) {
  if (variant == 1 || variant == 2 || variant == 3) return 0;
  long long lpad = (p->Lk + 7) / 8 * 8;
  return (size_t)p->B * p->H * HD * lpad * 2;
}

extern "C" int evo_attn_fwd_ws(const evo_attn_params* p, int variant, void* workspace, size_t workspace_bytes, void* stream) {
  EVO_REQUIRE(p->hd == HD, "evo_attn_fwd: head_dim %d unsupported (kernel is specialised for 128)", p->hd);
  EVO_REQUIRE(p->q_tok_stride % 8 == 0 && p->kv_tok_stride % 8 == 0 && p->q_batch_stride % 8 == 0 && p->kv_batch_stride % 8 == 0,
              "evo_attn_fwd: strides must be multiples of 8 elements");
  EVO_REQUIRE(p->q_pos0 + p->Lq <= p->Lk, "evo_attn_fwd: queries extend past the keys (q_pos0 %lld + Lq %lld > Lk %lld)",
              (long long)p->q_pos0, (long long)p->Lq, (long long)p->Lk);
  EVO_REQUIRE(p->n_out_peers == 0 || variant == 2 || variant == 3, "evo_attn_fwd: peer-scattered output is implemented by variants 2 and 3 only");
  if (p->Lq == 0 || p->B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  {
    uint64_t dims[4] = {(uint64_t)HD, (uint64_t)p->H, (uint64_t)p->Lq, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)HD * 2, (uint64_t)p->q_tok_stride * 2, (uint64_t)p->q_batch_stride * 2};
    uint32_t box[4] = {64, 1, BQ, 1};
    if ((rc = make_tmap_4d(&tmQ, p->q, dims, str, box))) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)HD, (uint64_t)p->H, (uint64_t)p->Lk, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)HD * 2, (uint64_t)p->kv_tok_stride * 2, (uint64_t)p->kv_batch_stride * 2};
    uint32_t box[4] = {64, 1, BKV, 1};
    if ((rc = make_tmap_4d(&tmK, p->k, dims, str, box))) return rc;
    if (variant == 1 || variant == 2 || variant == 3) { if ((rc = make_tmap_4d(&tmV, p->v, dims, str, box))) return rc; }
  }
  if (variant == 2 || variant == 3) return evo_attn_pp_launch(tmQ, tmK, tmV, p, st, variant == 3 ? 3 : 0);
  if (variant != 1) {
    long long lpad = (p->Lk + 7) / 8 * 8;
    size_t need = evo_attn_fwd_workspace(p, variant);
    EVO_REQUIRE(workspace && workspace_bytes >= need, "evo_attn_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    dim3 grid((unsigned)((lpad + 63) / 64), p->H, p->B);
    transpose_v_kernel<<<grid, 256, 0, st>>>((const bf16*)p->v, (bf16*)workspace, p->Lk, lpad, p->H, p->kv_tok_stride, p->kv_batch_stride);
    if ((rc = check_launch("transpose_v"))) return rc;
    uint64_t dims[4] = {(uint64_t)lpad, (uint64_t)HD, (uint64_t)p->H, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)lpad * 2, (uint64_t)lpad * HD * 2, (uint64_t)lpad * HD * p->H * 2};
    uint32_t box[4] = {64, HD, 1, 1};
    if ((rc = make_tmap_4d(&tmV, workspace, dims, str, box))) return rc;
  }
  AttArgs a;
  a.out = (bf16*)p->out; a.B = p->B; a.H = p->H; a.Lq = p->Lq; a.Lk = p->Lk; a.q_pos0 = p->q_pos0;
  a.scale_log2 = p->softmax_scale * 1.4426950408889634f;
  dim3 grid((unsigned)((p->Lq + BQ - 1) / BQ), p->H, p->B);
  if (variant == 1) {
    static unsigned long long done = 0;
    if ((rc = ensure_dyn_smem(attn_fwd_kernel<true>, ATT_SMEM, done))) return rc;
    attn_fwd_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, a);
  } else {
    static unsigned long long done = 0;
    if ((rc = ensure_dyn_smem(attn_fwd_kernel<false>, ATT_SMEM, done))) return rc;
    attn_fwd_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, a);
  }
  return check_launch("evo_attn_fwd");
}
