// Parquet column-chunk decode on the GPU.  The host frames the pages (parquet_meta.cc: Thrift headers,
// Snappy, run tables); these kernels expand a whole column chunk — any mix of PLAIN and dictionary-encoded pages — in one pass:
// every row finds its run by binary search over the chunk's run table (a few hundred entries, L1-resident), unpacks its
// dictionary index or reads its PLAIN value, and writes the Arrow value (narrowing INT32 -> int8 / int16, sign-extending
// INT32 / INT64 / big-endian FIXED_LEN_BYTE_ARRAY decimals to 128 bits).  Definition levels (max level 1) become the validity
// bytes; an exclusive scan of them gives every non-NULL row the ordinal of its stored value.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "kernels_parquet.cuh"

namespace b200q {

namespace {

constexpr int PB = 256;
int pgrid(int64_t n) {
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + PB * 4 - 1) / (PB * 4), (int64_t)sms * 8));
}

__device__ __forceinline__ int find_run(const PqDevRun* __restrict__ runs, int n, uint32_t pos) {
  int lo = 0, hi = n - 1;                                             // last run whose start <= pos
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (runs[mid].start <= pos) lo = mid; else hi = mid - 1; }
  return lo;
}
__device__ __forceinline__ unsigned long long read_bits(const uint8_t* __restrict__ bytes, unsigned long long bit, int bw) {
  if (bw == 0) return 0;
  const unsigned long long byte = bit >> 3; const int sh = (int)(bit & 7);
  unsigned long long w = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) w |= (unsigned long long)bytes[byte + i] << (8 * i);      // bw <= 32 and sh <= 7: 40 bits suffice (buffers are padded)
  return (w >> sh) & ((1ull << bw) - 1);
}

__global__ void __launch_bounds__(PB) pq_levels_kernel(const uint8_t* __restrict__ bytes, const PqDevRun* __restrict__ runs, int n_runs, long long n_rows, uint8_t* __restrict__ valid) {
  for (long long r = blockIdx.x * (long long)PB + threadIdx.x; r < n_rows; r += (long long)gridDim.x * PB) {
    const PqDevRun run = runs[find_run(runs, n_runs, (uint32_t)r)];
    valid[r] = run.kind == PQR_RLE ? (uint8_t)(run.off_or_value != 0) : (uint8_t)read_bits(bytes, run.off_or_value + (unsigned long long)((uint32_t)r - run.start), 1);
  }
}

__global__ void __launch_bounds__(PB) pq_decode_kernel(const PqDecodeSpec sp, const uint8_t* __restrict__ valid, const int32_t* __restrict__ ordinal, long long n_rows, void* out, int* err) {
  for (long long r = blockIdx.x * (long long)PB + threadIdx.x; r < n_rows; r += (long long)gridDim.x * PB) {
    const bool ok = valid ? valid[r] != 0 : true;
    unsigned long long lo = 0, hi = 0;                                // the stored value (low / high 64 bits)
    if (ok) {
      const uint32_t o = ordinal ? (uint32_t)ordinal[r] : (uint32_t)r;
      const PqDevRun run = sp.value_runs[find_run(sp.value_runs, sp.n_value_runs, o)];
      const uint32_t k = o - run.start;
      const uint8_t* src;
      if (run.kind == PQR_PLAIN) {
        if (sp.src_width == 0) { const unsigned long long bit = run.off_or_value * 8 + k; lo = (sp.bytes[bit >> 3] >> (bit & 7)) & 1; src = nullptr; }
        else src = sp.bytes + run.off_or_value + (unsigned long long)k * sp.src_width;
      } else {
        const unsigned long long idx = run.kind == PQR_RLE ? run.off_or_value : read_bits(sp.bytes, run.off_or_value + (unsigned long long)k * run.bw, run.bw);
        if (sp.src_width == 0) { lo = idx & 1; src = nullptr; }                                     // RLE-encoded Booleans (data page v2)
        else if (idx >= (unsigned long long)sp.dict_count) { atomicOr(err, 1); src = nullptr; }
        else src = sp.dict + idx * sp.src_width;
      }
      if (src) {
        if (sp.out_kind == PQO_DEC_FROM_FLBA) {                       // big-endian two's complement of src_width bytes
          const bool neg = src[0] & 0x80;
          lo = hi = neg ? ~0ull : 0ull;
          for (int i = 0; i < sp.src_width; i++) { hi = (hi << 8) | (lo >> 56); lo = (lo << 8) | src[i]; }
        } else if (sp.src_width == 4) { uint32_t v = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24); lo = (unsigned long long)(long long)(int32_t)v; hi = (long long)lo < 0 ? ~0ull : 0; }
        else { for (int i = 0; i < 8; i++) lo |= (unsigned long long)src[i] << (8 * i); hi = (long long)lo < 0 ? ~0ull : 0; }
      }
    }
    switch (sp.out_kind) {
      case PQO_I8: ((int8_t*)out)[r] = (int8_t)lo; break;
      case PQO_I16: ((int16_t*)out)[r] = (int16_t)lo; break;
      case PQO_I32: ((int32_t*)out)[r] = (int32_t)lo; break;
      case PQO_I64: ((unsigned long long*)out)[r] = lo; break;
      case PQO_BOOL_BYTES: ((uint8_t*)out)[r] = (uint8_t)(lo & 1); break;
      default: ((unsigned long long*)out)[2 * r] = lo; ((unsigned long long*)out)[2 * r + 1] = hi; break;
    }
  }
}

}  // namespace

int launch_pq_levels(const uint8_t* bytes, const PqDevRun* runs, int n_runs, int64_t n_rows, uint8_t* valid, cudaStream_t s) {
  if (n_rows <= 0) return 0;
  pq_levels_kernel<<<pgrid(n_rows), PB, 0, s>>>(bytes, runs, n_runs, n_rows, valid);
  return 1;
}
int launch_pq_decode(const PqDecodeSpec& sp, const uint8_t* valid, const int32_t* ordinal, int64_t n_rows, void* out, int* d_err, cudaStream_t s) {
  if (n_rows <= 0) return 0;
  pq_decode_kernel<<<pgrid(n_rows), PB, 0, s>>>(sp, valid, ordinal, n_rows, out, d_err);
  return 1;
}

}  // namespace b200q
