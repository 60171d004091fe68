// Does each die's L2 fetch a line from DRAM on its own?  (ground truth for evo_b200/csrc/die_map.cu and the GEMM's die-aware
// rasterisation.)  For every SM s: flush L2, let the CTAs on SM 0 read a 16 MB buffer (DRAM -> L2), then let the CTAs on SM s read
// the same buffer.  Under ncu (--cache-control none, dram__bytes_read.sum per launch) the second read costs ~0 bytes if the line is
// served from wherever the first read left it, ~16 MB if SM s sits on the other die AND that die's L2 goes to DRAM by itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o die_probe die_probe.cu
//   ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:touch --csv --log-file out.csv ./die_probe
// Launch order in the log: for s = 0 .. n_sm-1: touch(0), touch(s).  Without ncu the program prints the time of the second read.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

__global__ void flush(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(1, 2, 3, 4);
}

__global__ void touch(const uint4* p, size_t n, unsigned target, unsigned* sink) {
  if (smid() != target) return;
  unsigned acc = 0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    uint4 v; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i) : "memory");
    acc += v.x ^ v.w;
  }
  if (acc == 0x12345u) *sink = acc;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int n_sm = prop.multiProcessorCount;
  const size_t bytes = 16u << 20, fbytes = 512u << 20;
  uint4 *buf, *fl; unsigned* sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&fl, fbytes); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int s = 0; s < n_sm; ++s) {
    flush<<<n_sm * 4, 256>>>(fl, fbytes / 16);
    touch<<<n_sm * 8, 256>>>(buf, bytes / 16, 0, sink);
    cudaEventRecord(e0);
    touch<<<n_sm * 8, 256>>>(buf, bytes / 16, (unsigned)s, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("sm %d second-read %.3f ms\n", s, ms);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
