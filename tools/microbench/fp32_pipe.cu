// FP32 pipe micro-benchmark for sm_100a: issue rate of FFMA / FFMA2 / FMUL2 / FADD2 / HFMA2.BF16 with register operands,
// at 1, 2 and 4 warps per SM sub-partition.  Prints cycles per warp-instruction per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipe fp32_pipe.cu && ./fp32_pipe
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define CHAINS 16
#define ITERS 512

template <int MODE>
__global__ void k(float* out, long long* cyc, const float* __restrict__ in) {
  // MODE 0: FFMA (3 distinct regs)  1: FFMA2 (3 distinct 64-bit regs)  2: FFMA2 with a shared multiplicand  3: FMUL2  4: FADD2
  // 5: HFMA2.BF16 (3 regs)  6: FFMA2 acc = a*acc + b (recurrence shape: accumulator is the multiplicand)  7: FFMA with shared multiplicand
  float2 acc[CHAINS], a[CHAINS], b[CHAINS];
  // operands come from memory: run-time values the compiler cannot rematerialise inside the timed loop
  const float* src = in + threadIdx.x * 8 * CHAINS;
  for (int i = 0; i < CHAINS; ++i) {
    acc[i] = make_float2(src[i * 8 + 0], src[i * 8 + 1]); a[i] = make_float2(src[i * 8 + 2], src[i * 8 + 3]); b[i] = make_float2(src[i * 8 + 4], src[i * 8 + 5]);
  }
  unsigned int au[CHAINS], bu[CHAINS], cu[CHAINS];
  for (int i = 0; i < CHAINS; ++i) { au[i] = __float_as_uint(src[i * 8 + 6]); bu[i] = __float_as_uint(src[i * 8 + 7]); cu[i] = __float_as_uint(src[i * 8 + 5]); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (MODE == 0) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a[i].x), "f"(b[i].x)); }
      else if (MODE == 7) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a[0].x), "f"(b[i].x)); }
      else if (MODE == 1) { unsigned long long &A = *(unsigned long long*)&acc[i], &X = *(unsigned long long*)&a[i], &Y = *(unsigned long long*)&b[i];
                            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A) : "l"(X), "l"(Y)); }
      else if (MODE == 2) { unsigned long long &A = *(unsigned long long*)&acc[i], &X = *(unsigned long long*)&a[0], &Y = *(unsigned long long*)&b[i];
                            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A) : "l"(X), "l"(Y)); }
      else if (MODE == 3) { unsigned long long &A = *(unsigned long long*)&acc[i], &X = *(unsigned long long*)&a[i];
                            asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A) : "l"(X)); }
      else if (MODE == 4) { unsigned long long &A = *(unsigned long long*)&acc[i], &X = *(unsigned long long*)&b[i];
                            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A) : "l"(X)); }
      else if (MODE == 5) { asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(cu[i]) : "r"(au[i]), "r"(bu[i])); }
      else if (MODE == 8) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(cu[i]) : "f"(acc[i].x), "f"(acc[i].y)); acc[i].x = __uint_as_float(cu[i]); }
      else if (MODE == 9) { asm volatile("shfl.sync.bfly.b32 %0, %0, 16, 0x1f, 0xffffffff;" : "+r"(cu[i])); }
      else if (MODE == 10) { asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(cu[i]) : "r"(au[i])); }
      else if (MODE == 6) { unsigned long long &A = *(unsigned long long*)&acc[i], &X = *(unsigned long long*)&a[i], &Y = *(unsigned long long*)&b[i];
                            asm volatile("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(A) : "l"(X), "l"(Y)); }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < CHAINS; ++i) s += acc[i].x + acc[i].y + __uint_as_float(cu[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name) {
  float* out; long long* cyc; float* in;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&in, 1024 * 8 * CHAINS * 4);
  { static float h[1024 * 8 * CHAINS]; for (int i = 0; i < 1024 * 8 * CHAINS; ++i) h[i] = 1.0f + 1e-6f * (i % 97) - 5e-5f; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice); }
  for (int warps_per_smsp : {1, 2, 4}) {
    int threads = warps_per_smsp * 4 * 32;
    k<MODE><<<148, threads>>>(out, cyc, in);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, cyc, in);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double per = avg / ((double)ITERS * CHAINS * warps_per_smsp);
    printf("{\"what\": \"fp32_pipe\", \"op\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_warp_instr_per_smsp\": %.3f}\n", name, warps_per_smsp, per);
  }
  cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
  run<0>("FFMA r,r,r"); run<7>("FFMA shared-a"); run<1>("FFMA2 r,r,r"); run<2>("FFMA2 shared-a"); run<6>("FFMA2 acc=a*acc+b"); run<3>("FMUL2"); run<4>("FADD2"); run<5>("HFMA2.BF16"); run<8>("F2FP.BF16.PACK + MOV"); run<9>("SHFL.BFLY"); run<10>("PRMT");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
