// TEST COMPARATOR ONLY: plain cuBLASLt GEMM (C = A.W^T [+ bias]) used by tests/ to check
// the tcgen05 kernel on the GPU.  Never called from the product path.
#include "common.cuh"
#include "../../include/evo_b200.h"
#include <cublasLt.h>

using namespace evo;

#define EVO_LT(call)                                                                  \
  do { cublasStatus_t s_ = (call); if (s_ != CUBLAS_STATUS_SUCCESS) {                 \
    ::evo::set_error("%s failed: cublas status %d", #call, (int)s_); return -4; } } while (0)

extern "C" int evo_gemm_cublaslt_reference(const evo_gemm_params* p, void* workspace, size_t workspace_bytes, void* stream) {
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
    set_error("evo_gemm_cublaslt_reference: only NONE/BIAS epilogues");
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
  EVO_REQUIRE(found > 0, "evo_gemm_cublaslt_reference: no algorithm");
  float alpha = 1.f, beta = 0.f;
  EVO_LT(cublasLtMatmul(handle, op, &alpha, p->W, la, p->A, lb, &beta, p->C, lc, p->C, lc, &heur.algo, workspace, workspace_bytes, (cudaStream_t)stream));
  cublasLtMatmulPreferenceDestroy(pref);
  cublasLtMatrixLayoutDestroy(la); cublasLtMatrixLayoutDestroy(lb); cublasLtMatrixLayoutDestroy(lc);
  cublasLtMatmulDescDestroy(op);
  return 0;
}
