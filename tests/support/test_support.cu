// Test-support library (tests/support/libevo_b200_test.so) -- NOT part of the product.
// Comparators the GPU tests check the product kernels against, kept out of libevo_b200.so and out
// of include/evo_b200.h: a cuBLASLt GEMM, a one-warp-per-row CUDA-core
// attention and a bf16 add.  Built by evo_b200/build.py:build_test_support(); links cuBLASLt
// (the product library does not).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cublasLt.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/evo_b200.h"

typedef __nv_bfloat16 bf16;
static thread_local char t_err[512] = "";
static void tset_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(t_err, sizeof(t_err), fmt, ap); va_end(ap); }
#define T_REQUIRE(cond, ...) do { if (!(cond)) { tset_error(__VA_ARGS__); return -1; } } while (0)
static int t_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { tset_error("%s: launch failed: %s", what, cudaGetErrorString(e)); return -3; }
  return 0;
}
extern "C" const char* evot_last_error(void) { return t_err; }

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
constexpr int HD = 128;

// ---- plain CUDA-core causal attention: one warp per query row, online softmax in fp32
__global__ void attn_simple_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, bf16* __restrict__ out,
                                   long long q_tok, long long kv_tok, long long q_batch, long long kv_batch,
                                   int B, long long Lq, long long Lk, int H, long long q_pos0, float scale) {
  __shared__ float qs[4][HD];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * 4 + w;         // (b, h, i)
  const bool active = row < (long long)B * H * Lq;
  long long i = active ? row % Lq : 0;
  int h = active ? (int)((row / Lq) % H) : 0;
  int b = active ? (int)(row / (Lq * H)) : 0;
  const bf16* qp = q + b * q_batch + i * q_tok + (long long)h * HD;
  for (int d = lane; d < HD; d += 32) qs[w][d] = __bfloat162float(qp[d]);
  __syncwarp();
  if (!active) return;
  const long long pos = q_pos0 + i;
  const long long nk = min(Lk, pos + 1);
  float m = -INFINITY, l = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long j0 = 0; j0 < nk; j0 += 32) {
    long long j = j0 + lane;
    float s = -INFINITY;
    if (j < nk) {
      const bf16* kp = k + b * kv_batch + j * kv_tok + (long long)h * HD;
      float acc = 0.f;
      for (int d = 0; d < HD; ++d) acc = fmaf(qs[w][d], __bfloat162float(kp[d]), acc);
      s = acc * scale;
    }
    float mc = s;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mc = fmaxf(mc, __shfl_xor_sync(0xffffffffu, mc, off));
    float mn = fmaxf(m, mc);
    float alpha = (m == -INFINITY) ? 0.f : expf(m - mn);
    float p = (j < nk) ? expf(s - mn) : 0.f;
    float ps = p;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
    l = l * alpha + ps;
#pragma unroll
    for (int t = 0; t < 4; ++t) o[t] *= alpha;
    float pb = rbf(p);
    for (int jj = 0; jj < 32; ++jj) {
      float pj = __shfl_sync(0xffffffffu, pb, jj);
      if (j0 + jj < nk) {
        const bf16* vp = v + b * kv_batch + (j0 + jj) * kv_tok + (long long)h * HD;
#pragma unroll
        for (int t = 0; t < 4; ++t) o[t] = fmaf(pj, __bfloat162float(vp[lane + 32 * t]), o[t]);
      }
    }
    m = mn;
  }
  bf16* op = out + ((b * Lq + i) * H + h) * HD;
#pragma unroll
  for (int t = 0; t < 4; ++t) op[lane + 32 * t] = __float2bfloat16_rn(o[t] / l);
}


extern "C" int evot_attn_fwd_simple(const evo_attn_params* p, void* stream) {
  T_REQUIRE(p->hd == HD, "evot_attn_fwd_simple: head_dim %d unsupported", p->hd);
  long long rows = (long long)p->B * p->H * p->Lq;
  if (rows == 0) return 0;
  attn_simple_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
      (const bf16*)p->q, (const bf16*)p->k, (const bf16*)p->v, (bf16*)p->out, p->q_tok_stride, p->kv_tok_stride,
      p->q_batch_stride, p->kv_batch_stride, p->B, p->Lq, p->Lk, p->H, p->q_pos0, p->softmax_scale);
  return t_check_launch("evot_attn_fwd_simple");
}

// ---- bf16 add (fp32 sum, one rounding)
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}
extern "C" int evot_add(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n);
  return t_check_launch("evot_add");
}

// ---- cuBLASLt GEMM: C = A.W^T [+ bias]
#define EVO_LT(call)                                                                  \
  do { cublasStatus_t s_ = (call); if (s_ != CUBLAS_STATUS_SUCCESS) {                 \
    tset_error("%s failed: cublas status %d", #call, (int)s_); return -4; } } while (0)

extern "C" int evot_gemm_cublaslt(const evo_gemm_params* p, void* workspace, size_t workspace_bytes, void* stream) {
  static cublasLtHandle_t handle = nullptr;
  if (!handle) EVO_LT(cublasLtCreate(&handle));
  // row-major C[M,N] = A[M,K] W[N,K]^T  <=>  column-major C^T[N,M] = W^T... : op(W)=T (K x N col-major view), op(A)=N
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr;
  cublasLtMatmulPreference_t pref = nullptr;
  EVO_LT(cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
  cublasOperation_t tA = CUBLAS_OP_T, tB = CUBLAS_OP_N;
  EVO_LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &tA, sizeof(tA)));
  EVO_LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &tB, sizeof(tB)));
  if (p->epilogue == EVO_EPI_BIAS) {
    cublasLtEpilogue_t ep = CUBLASLT_EPILOGUE_BIAS;
    EVO_LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)));
    EVO_LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &p->bias, sizeof(p->bias)));
    cudaDataType_t bt = CUDA_R_16BF;
    EVO_LT(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &bt, sizeof(bt)));
  } else if (p->epilogue != EVO_EPI_NONE) {
    tset_error("evot_gemm_cublaslt: only NONE/BIAS epilogues");
    return -1;
  }
  // "A" of cuBLAS = W stored (K x N) column-major with ld K; "B" = A stored (K x M) column-major with ld lda
  EVO_LT(cublasLtMatrixLayoutCreate(&la, CUDA_R_16BF, p->K, p->N, p->K));
  EVO_LT(cublasLtMatrixLayoutCreate(&lb, CUDA_R_16BF, p->K, p->M, p->lda));
  EVO_LT(cublasLtMatrixLayoutCreate(&lc, CUDA_R_16BF, p->N, p->M, p->ldc));
  EVO_LT(cublasLtMatmulPreferenceCreate(&pref));
  EVO_LT(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &workspace_bytes, sizeof(workspace_bytes)));
  cublasLtMatmulHeuristicResult_t heur;
  int found = 0;
  EVO_LT(cublasLtMatmulAlgoGetHeuristic(handle, op, la, lb, lc, lc, pref, 1, &heur, &found));
  T_REQUIRE(found > 0, "evot_gemm_cublaslt: no algorithm");
  float alpha = 1.f, beta = 0.f;
  EVO_LT(cublasLtMatmul(handle, op, &alpha, p->W, la, p->A, lb, &beta, p->C, lc, p->C, lc, &heur.algo, workspace, workspace_bytes, (cudaStream_t)stream));
  cublasLtMatmulPreferenceDestroy(pref);
  cublasLtMatrixLayoutDestroy(la); cublasLtMatrixLayoutDestroy(lb); cublasLtMatrixLayoutDestroy(lc);
  cublasLtMatmulDescDestroy(op);
  return 0;
}
