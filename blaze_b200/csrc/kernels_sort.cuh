// Descriptors and launchers of the SortExec kernels (kernels_sort.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace b200q {

struct SortKeyCol {
  const void* values;                 // contiguous column of the concatenated input
  const uint8_t* valid_bytes;         // one byte per row, null: no NULLs
  uint8_t phys;                       // PhysKind
  uint8_t descending, nulls_first;
  uint8_t dec_word;                   // decimal128: 0 = low word, 1 = high word
  uint32_t _pad;
  unsigned long long mask;            // all-ones over the type's width (keeps `~w` of a descending key inside it)
};

int launch_sort_iota(uint32_t* d_idx, int64_t n, cudaStream_t s);
int launch_sort_normalise(const SortKeyCol& k, const uint32_t* d_idx, int64_t n, unsigned long long* d_keys, uint8_t* d_nullrank, cudaStream_t s);
// d_hist: 9 x 256 counters (zeroed): digits 0..7 of the keys, then the NULL ranks
int launch_sort_digit_hist(const unsigned long long* d_keys, const uint8_t* d_nullrank, int64_t n, unsigned long long* d_hist, cudaStream_t s);
int64_t sort_num_tiles(int64_t n);
// one stable pass on digit `shift / 8` (shift < 0: on the NULL rank): (keys, nullrank, idx) -> (okeys, onull, oidx)
// d_counts / d_offs: 256 * sort_num_tiles(n) + 1 int32 each, d_block_sums: scan_num_blocks(256 * tiles) int32
int launch_sort_pass(const unsigned long long* d_keys, const uint8_t* d_nullrank, const uint32_t* d_idx, int64_t n, int shift, int32_t* d_counts, int32_t* d_offs, int32_t* d_block_sums,
                     unsigned long long* d_okeys, uint8_t* d_onull, uint32_t* d_oidx, cudaStream_t s);

}  // namespace b200q
