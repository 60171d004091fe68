// Batch front-end on the device: raw sequence bytes -> padded token-id matrix.
//
// The reference's prepare_batch (evo/scoring.py:9-33) tokenises on the host into Python lists, builds one int64 tensor per
// sequence and copies each to the GPU separately.  Here the host ships the concatenated bytes once (1 byte per
// nucleotide instead of 8) plus B+1 offsets, and this kernel lays out (B, width) ids: [BOS] + bytes + pad_id...
// CharLevelTokenizer.tokenize is the identity on bytes (evo/tokenizer.py:41), so no table is needed.
#include "common.cuh"
#include "../../include/evo_b200.h"

using namespace evo;

namespace {

template <typename T>
__global__ void tokenize_pad_kernel(const uint8_t* __restrict__ bytes, const long long* __restrict__ offsets, T* __restrict__ ids,
                                    int B, long long width, int bos, int bos_id, int pad_id) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= width) return;
  const long long o0 = offsets[b], len = offsets[b + 1] - o0;
  T v;
  if (j < bos) v = (T)bos_id;
  else if (j - bos < len) v = (T)bytes[o0 + j - bos];
  else v = (T)pad_id;
  ids[(long long)b * width + j] = v;
}

}  // namespace

extern "C" int evo_tokenize_pad(const void* bytes, const int64_t* offsets, void* ids_out, int ids_are_i64, int B, int64_t width,
                                int prepend_bos, int bos_id, int pad_id, void* stream) {
  EVO_REQUIRE(B >= 0 && B <= 65535 && width >= 0, "evo_tokenize_pad: bad shape (B=%d, width=%lld)", B, (long long)width);
  if (B == 0 || width == 0) return 0;
  dim3 grid((unsigned)((width + 255) / 256), B);
  if (ids_are_i64) tokenize_pad_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)bytes, (const long long*)offsets, (long long*)ids_out, B, width, prepend_bos ? 1 : 0, bos_id, pad_id);
  else tokenize_pad_kernel<int><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)bytes, (const long long*)offsets, (int*)ids_out, B, width, prepend_bos ? 1 : 0, bos_id, pad_id);
  return check_launch("evo_tokenize_pad");
}
