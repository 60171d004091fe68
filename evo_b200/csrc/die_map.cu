// Which of the two dies does an SM sit on?  (used by the big-tile GEMM's die-aware rasterisation, gemm_tcgen05.cu)
//
// B200 is two dies, each with half of the SMs and half of the L2; an L2 hit on the SM's own die costs ~234 cycles, on
// the other die ~262 (B300_MICROARCH "SM->L2-die routing").  The SM -> die map differs per physical GPU, so it is
// measured once per device: one thread per SM, one SM at a time (a ticket lock, so the probes do not queue behind each
// other), chases a pointer that points at itself through L2 (ld.global.cg) on P lines spread over a buffer; for every
// line the SMs fall into a near and a far group, and two SMs are on the same die iff they are in the same group on
// (nearly) every line.  A second kernel checks that the two CTAs of a 2-CTA cluster land on the SMs 2t, 2t+1 of one TPC,
// which is what lets a CTA pair derive its per-die slot from %smid alone.
// The result is only used if it is self-consistent (both dies populated with an even SM count, TPC-mates together, every
// SM decided with a clear majority); otherwise the GEMM keeps its die-oblivious tile order.
#include "common.cuh"
#include "die_classify.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

using namespace evo;

namespace {

constexpr int MAX_SM = 256, PROBES = 24, REPS = 48, LINE_STRIDE = 6 * 1024 + 128;   // bytes between probe lines

__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long ld_cg(unsigned long long addr) {
  unsigned long long v; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory"); return v;
}

__global__ void die_chase_init(unsigned long long* buf) {
  if (threadIdx.x < PROBES) { unsigned long long* p = buf + (size_t)threadIdx.x * LINE_STRIDE / 8; *p = (unsigned long long)p; }
}

// ctl[0] = tickets handed out, ctl[1] = ticket being served; claimed[sm] = an earlier CTA of this SM took the job
__global__ void die_latency_kernel(const unsigned long long* buf, unsigned* claimed, unsigned* ctl, float* lat) {
  if (threadIdx.x != 0) return;
  const unsigned sm = smid();
  if (sm >= MAX_SM || atomicCAS(&claimed[sm], 0u, 1u) != 0u) return;
  const unsigned my = atomicAdd(&ctl[0], 1u);
  while (*(volatile unsigned*)&ctl[1] != my) __nanosleep(400);
  unsigned long long sink = 0;
  for (int p = 0; p < PROBES; ++p) {
    unsigned long long x = (unsigned long long)(buf + (size_t)p * LINE_STRIDE / 8);
    for (int i = 0; i < 4; ++i) x = ld_cg(x);                       // bring the line into L2, warm the TLB
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REPS; ++i) x = ld_cg(x);
    const long long t1 = clock64();
    lat[sm * PROBES + p] = (float)(t1 - t0) / REPS;
    sink += x;
  }
  if (sink == 1) lat[0] = 0.f;                                       // keeps the chain alive
  __threadfence();
  atomicAdd(&ctl[1], 1u);
}

// one 2-CTA cluster per TPC (the GEMM's launch shape): where do the two CTAs of a cluster land?
__global__ void pair_probe_kernel(unsigned* where) {
  if (threadIdx.x == 0) { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); where[blockIdx.x] = smid() | (r << 16); }
}

struct Entry { std::once_flag once; DieMap map; std::atomic<unsigned> next_line{0}; };
Entry g_entries[64];

void dump(const char* path, int sms, const std::vector<float>& lat, const std::vector<int>& die, const char* verdict) {
  FILE* f = fopen(path, "w");
  if (!f) return;
  fprintf(f, "# %s\n# sm die latency[%d probes] (cycles per dependent L2 load)\n", verdict, PROBES);
  for (int s = 0; s < sms; ++s) {
    fprintf(f, "%d %d", s, die[s]);
    for (int p = 0; p < PROBES; ++p) fprintf(f, " %.1f", lat[s * PROBES + p]);
    fprintf(f, "\n");
  }
  fclose(f);
}

bool calibrate(DieMap& out) {
  const int sms = device_sm_count();
  if (sms <= 0 || sms > MAX_SM || (sms & 1)) return false;
  cudaStream_t st = nullptr;
  unsigned long long* buf = nullptr; unsigned* ctl = nullptr; float* lat_d = nullptr; unsigned* where_d = nullptr;
  const size_t buf_bytes = (size_t)PROBES * LINE_STRIDE + 256;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&buf, buf_bytes) == cudaSuccess && cudaMalloc(&ctl, (MAX_SM + 2) * sizeof(unsigned)) == cudaSuccess &&
       cudaMalloc(&lat_d, MAX_SM * PROBES * sizeof(float)) == cudaSuccess && cudaMalloc(&where_d, MAX_SM * sizeof(unsigned)) == cudaSuccess;
  std::vector<float> lat(MAX_SM * PROBES, 0.f);
  std::vector<unsigned> claimed(MAX_SM + 2, 0u), where(MAX_SM, 0u);
  if (ok) {
    cudaMemsetAsync(ctl, 0, (MAX_SM + 2) * sizeof(unsigned), st);
    cudaMemsetAsync(lat_d, 0, MAX_SM * PROBES * sizeof(float), st);
    die_chase_init<<<1, 32, 0, st>>>(buf);
    die_latency_kernel<<<sms * 16, 32, 0, st>>>(buf, ctl + 2, ctl, lat_d);        // first pass: brings the clocks up; discarded
    cudaMemsetAsync(ctl, 0, (MAX_SM + 2) * sizeof(unsigned), st);
    die_latency_kernel<<<sms * 16, 32, 0, st>>>(buf, ctl + 2, ctl, lat_d);
    // the pair probe mirrors the GEMM: sms CTAs in clusters of two, enough dynamic shared memory for one CTA per SM
    const int smem = 160 * 1024;
    ok = cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
    if (ok) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(sms); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      ok = cudaLaunchKernelEx(&cfg, pair_probe_kernel, where_d) == cudaSuccess;
    }
    ok = ok && cudaMemcpyAsync(lat.data(), lat_d, MAX_SM * PROBES * sizeof(float), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(claimed.data(), ctl, (MAX_SM + 2) * sizeof(unsigned), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(where.data(), where_d, MAX_SM * sizeof(unsigned), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
  }
  cudaFree(buf); cudaFree(ctl); cudaFree(lat_d); cudaFree(where_d);
  if (st) cudaStreamDestroy(st);
  if (!ok) { cudaGetLastError(); return false; }

  const char* verdict = "ok";
  std::vector<int> die(sms, -1);
  // every SM must have been measured
  for (int s = 0; s < sms && ok; ++s) if (!claimed[2 + s]) { ok = false; verdict = "an SM was never measured"; }
  if (ok) { const char* why = classify_dies(lat, sms, PROBES, where, die); if (why) { ok = false; verdict = why; } }
  int count[2] = {0, 0};
  if (ok) for (int s = 0; s < sms; ++s) ++count[die[s]];
  if (const char* path = getenv("EVO_B200_GEMM_DIE_DUMP")) dump(path, sms, lat, die, ok ? "ok" : verdict);
  if (!ok) return false;

  std::vector<uint16_t> tab(MAX_SM / 2, 0);
  int next[2] = {0, 0};
  for (int t = 0; t < sms / 2; ++t) { const int d = die[2 * t]; tab[t] = (uint16_t)(d | (next[d]++ << 1)); }
  uint16_t* tab_d = nullptr;
  if (cudaMalloc(&tab_d, tab.size() * sizeof(uint16_t)) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaMemcpy(tab_d, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(tab_d); cudaGetLastError(); return false; }
  unsigned* claims = nullptr;
  if (next[0] + next[1] > DieMap::CLAIM_WORDS || cudaMalloc(&claims, (size_t)DieMap::CLAIM_LINES * DieMap::CLAIM_WORDS * sizeof(unsigned)) != cudaSuccess) {
    cudaFree(tab_d); cudaGetLastError(); return false;
  }
  out.valid = true; out.pairs[0] = next[0]; out.pairs[1] = next[1]; out.tab = tab_d; out.claims = claims;
  return true;
}

}  // namespace

const DieMap* evo::die_map(cudaStream_t caller) {
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return nullptr;
  Entry& e = g_entries[dev];
  if (caller) {      // never calibrate from inside a stream capture (the calibration synchronises); the caller falls back
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(caller, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return e.map.valid ? &e.map : nullptr;
    }
  }
  std::call_once(e.once, [&] { e.map = DieMap{}; calibrate(e.map); });
  return e.map.valid ? &e.map : nullptr;
}

unsigned* evo::die_next_claims(const DieMap* dm) {
  for (Entry& e : g_entries)
    if (&e.map == dm) return dm->claims + (size_t)(e.next_line.fetch_add(1) % DieMap::CLAIM_LINES) * DieMap::CLAIM_WORDS;
  return dm->claims;
}
