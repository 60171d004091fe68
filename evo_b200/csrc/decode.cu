// Decode-step kernels (one new token per sequence): everything a CUDA-graph replay of the
// step needs, with the sequence position read from DEVICE memory so the same captured graph
// serves every step (the reference re-launches ~20 tiny kernels per layer per token from
// Python).
#include "common.cuh"
#include "../../include/evo_b200.h"

using namespace evo;

namespace {

constexpr int HD = 128;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// out[m, g*128 + c] = bf16(gelu(t[m, g*256 + c])) * t[m, g*256 + 128 + c]   (t = [l1 | l2] interleaved GEMM output)
__global__ void gelu_gate_kernel(const bf16* __restrict__ t, bf16* __restrict__ out, long long M, int ipad) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (m, pair of columns)
  long long total = M * (ipad / 2);
  if (idx >= total) return;
  int c2 = (int)(idx % (ipad / 2)) * 2;
  long long m = idx / (ipad / 2);
  int g = c2 / 128, c = c2 % 128;
  const uint32_t a = *reinterpret_cast<const uint32_t*>(t + m * 2 * ipad + g * 256 + c);
  const uint32_t b = *reinterpret_cast<const uint32_t*>(t + m * 2 * ipad + g * 256 + 128 + c);
  float o0 = rbf(gelu_erf(bf_lo(a))) * bf_lo(b), o1 = rbf(gelu_erf(bf_hi(a))) * bf_hi(b);
  *reinterpret_cast<uint32_t*>(out + m * ipad + c2) = pack_bf16(o0, o1);
}

// L == 1: rotary on q,k at position *pos, k,v appended to the cache at row *pos.
// qkv (B, 3, H, 128) in place; cos/sin tables indexed by absolute position.
__global__ void decode_qkv_prep_kernel(bf16* __restrict__ qkv, bf16* __restrict__ cache, const bf16* __restrict__ cos, const bf16* __restrict__ sin,
                                       const long long* __restrict__ pos_ptr, int B, int H, long long max_seqlen) {
  pdl_launch_dependents(); pdl_wait();
  const long long pos = *pos_ptr;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (b, h, i) i in [0, 64)
  if (idx >= B * H * 64 || pos >= max_seqlen) return;
  int i = idx % 64, h = (idx / 64) % H, b = idx / (64 * H);
  const float c = __bfloat162float(cos[pos * 64 + i]), s = __bfloat162float(sin[pos * 64 + i]);
  bf16* q = qkv + ((long long)(b * 3 + 0) * H + h) * HD;
  bf16* k = qkv + ((long long)(b * 3 + 1) * H + h) * HD;
  const bf16* v = qkv + ((long long)(b * 3 + 2) * H + h) * HD;
  float q0 = __bfloat162float(q[i]), q1 = __bfloat162float(q[i + 64]);
  float k0 = __bfloat162float(k[i]), k1 = __bfloat162float(k[i + 64]);
  bf16 kr0 = __float2bfloat16_rn(k0 * c - k1 * s), kr1 = __float2bfloat16_rn(k0 * s + k1 * c);
  q[i] = __float2bfloat16_rn(q0 * c - q1 * s);
  q[i + 64] = __float2bfloat16_rn(q0 * s + q1 * c);
  k[i] = kr0; k[i + 64] = kr1;
  bf16* ck = cache + (((long long)b * max_seqlen + pos) * 2 + 0) * H * HD + (long long)h * HD;
  bf16* cv = ck + (long long)H * HD;
  ck[i] = kr0; ck[i + 64] = kr1;
  cv[i] = v[i]; cv[i + 64] = v[i + 64];
}

// Single-query attention over the KV cache, keys [0, *pos].  grid (H, B, nsplit); each CTA scans a
// contiguous key range: thread t owns keys t, t+128, ... of the range (one 256-byte K row and one
// V row per key), keeps an online-softmax state and a 128-wide fp32 accumulator in registers, and the
// 128 partial states are merged through shared memory; splits are merged by the last kernel.
constexpr int DT = 128;
__global__ void __launch_bounds__(DT) decode_attn_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ cache, float* __restrict__ part_o,
                                                         float* __restrict__ part_ml, const long long* __restrict__ pos_ptr,
                                                         int H, long long max_seqlen, int nsplit, float scale) {
  pdl_launch_dependents(); pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, tid = threadIdx.x;
  const long long nk = min(*pos_ptr + 1, max_seqlen);
  const long long per = (nk + nsplit - 1) / nsplit;
  const long long k0 = (long long)sp * per, k1 = min(nk, k0 + per);
  __shared__ float qs[HD];
  __shared__ float red_m[DT], red_l[DT];
  __shared__ float red_o[32][HD + 1];
  qs[tid] = __bfloat162float(qkv[((long long)(b * 3) * H + h) * HD + tid]) * scale;
  __syncthreads();
  float m = -INFINITY, l = 0.f, o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  const long long row_stride = 2LL * H * HD;
  const bf16* kb = cache + ((long long)b * max_seqlen) * row_stride + (long long)h * HD;
  for (long long j = k0 + tid; j < k1; j += DT) {
    const uint4* kr = reinterpret_cast<const uint4*>(kb + j * row_stride);
    const uint4* vr = reinterpret_cast<const uint4*>(kb + j * row_stride + (long long)H * HD);
    float sdot = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      uint4 kv = __ldg(kr + c);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&kv);
#pragma unroll
      for (int e = 0; e < 4; ++e) { sdot = fmaf(qs[c * 8 + 2 * e], bf_lo(w[e]), sdot); sdot = fmaf(qs[c * 8 + 2 * e + 1], bf_hi(w[e]), sdot); }
    }
    const float mn = fmaxf(m, sdot);
    const float alpha = __expf(m - mn), p = __expf(sdot - mn);
    const float pb = rbf(p);                      // P rounded to bf16 before the PV product, like the prefill kernel
    l = l * alpha + p;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      uint4 vv = __ldg(vr + c);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&vv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[c * 8 + 2 * e] = fmaf(pb, bf_lo(w[e]), o[c * 8 + 2 * e] * alpha);
        o[c * 8 + 2 * e + 1] = fmaf(pb, bf_hi(w[e]), o[c * 8 + 2 * e + 1] * alpha);
      }
    }
    m = mn;
  }
  // merge the 128 per-thread states: common max, rescale, then reduce 32 threads at a time through smem
  red_m[tid] = m;
  __syncthreads();
  float M = -INFINITY;
  for (int i = 0; i < DT; ++i) M = fmaxf(M, red_m[i]);
  const float sc = (m == -INFINITY) ? 0.f : __expf(m - M);
  red_l[tid] = l * sc;
  __syncthreads();
  float acc = 0.f;                                 // thread tid owns output dim tid
  for (int round = 0; round < DT / 32; ++round) {
    if (tid / 32 == round) {
#pragma unroll
      for (int d = 0; d < HD; ++d) red_o[tid % 32][d] = o[d] * sc;
    }
    __syncthreads();
    for (int i = 0; i < 32; ++i) acc += red_o[i][tid];
    __syncthreads();
  }
  float L = 0.f;
  for (int i = 0; i < DT; ++i) L += red_l[i];
  const long long pidx = ((long long)b * H + h) * nsplit + sp;
  part_o[pidx * HD + tid] = acc;
  if (tid == 0) { part_ml[pidx * 2] = M; part_ml[pidx * 2 + 1] = L; }
}

__global__ void decode_attn_merge_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, bf16* __restrict__ out,
                                         int H, int nsplit) {
  pdl_launch_dependents(); pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const long long base = ((long long)b * H + h) * nsplit;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, part_ml[(base + s) * 2]);
  float L = 0.f, acc = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = part_ml[(base + s) * 2];
    const float w = (ms == -INFINITY) ? 0.f : __expf(ms - M);
    L += part_ml[(base + s) * 2 + 1] * w;
    acc += part_o[(base + s) * HD + d] * w;
  }
  out[((long long)b * H + h) * HD + d] = __float2bfloat16_rn(acc / L);
}

__global__ void add_i64_kernel(long long* p, long long v) { *p += v; }

}  // namespace

extern "C" int evo_gelu_gate_interleaved(const void* t, void* out, int64_t M, int ipad, void* stream) {
  EVO_REQUIRE(ipad % 128 == 0, "evo_gelu_gate_interleaved: ipad must be a multiple of 128");
  long long total = M * (ipad / 2);
  if (total == 0) return 0;
  gelu_gate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)t, (bf16*)out, M, ipad);
  return check_launch("evo_gelu_gate_interleaved");
}

extern "C" int evo_decode_qkv_prep(void* qkv, void* cache, const void* cos, const void* sin, const int64_t* pos,
                                   int B, int H, int hd, int64_t max_seqlen, void* stream) {
  EVO_REQUIRE(hd == HD, "evo_decode_qkv_prep: head_dim %d unsupported", hd);
  int n = B * H * 64;
  EVO_CUDA(launch_pdl(decode_qkv_prep_kernel, dim3((n + 127) / 128), dim3(128), 0, (cudaStream_t)stream, (bf16*)qkv, (bf16*)cache, (const bf16*)cos, (const bf16*)sin,
                      (const long long*)pos, B, H, max_seqlen));
  return check_launch("evo_decode_qkv_prep");
}

extern "C" size_t evo_decode_attn_workspace(int B, int H, int nsplit) { return (size_t)B * H * nsplit * (HD + 2) * sizeof(float); }

extern "C" int evo_decode_attn(const void* qkv, const void* cache, void* out, const int64_t* pos, int B, int H, int hd,
                               int64_t max_seqlen, int nsplit, float softmax_scale, void* workspace, size_t workspace_bytes, void* stream) {
  EVO_REQUIRE(hd == HD, "evo_decode_attn: head_dim %d unsupported", hd);
  EVO_REQUIRE(nsplit >= 1 && nsplit <= 64, "evo_decode_attn: bad nsplit %d", nsplit);
  EVO_REQUIRE(workspace && workspace_bytes >= evo_decode_attn_workspace(B, H, nsplit), "evo_decode_attn: workspace too small");
  float* part_o = (float*)workspace;
  float* part_ml = part_o + (size_t)B * H * nsplit * HD;
  EVO_CUDA(launch_pdl(decode_attn_kernel, dim3(H, B, nsplit), dim3(DT), 0, (cudaStream_t)stream, (const bf16*)qkv, (const bf16*)cache, part_o, part_ml,
                      (const long long*)pos, H, max_seqlen, nsplit, softmax_scale));
  int rc = check_launch("evo_decode_attn");
  if (rc) return rc;
  EVO_CUDA(launch_pdl(decode_attn_merge_kernel, dim3(H, B), dim3(HD), 0, (cudaStream_t)stream, (const float*)part_o, (const float*)part_ml, (bf16*)out, H, nsplit));
  return check_launch("evo_decode_attn_merge");
}

extern "C" int evo_advance_position(int64_t* pos, int64_t delta, void* stream) {
  add_i64_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)pos, delta);
  return check_launch("evo_advance_position");
}
