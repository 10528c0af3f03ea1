// Descriptors and launchers of the ShuffleWriterExec kernels (kernels_shuffle.cu): hash partition + batch_serde encode.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace b200q {

constexpr int SHUF_MAX_COLS = 32;
constexpr int SHUF_MAX_PARTS = 4096;        // partition counters live in shared memory
constexpr int SHUF_TILE = 4096;             // rows ranked and staged per CTA step

// One column of the batch being written.  Wire layout of a record of m rows (datafusion-ext-commons/src/io/batch_serde.rs:66-77,
// 264-306, 530-551): varint(m), then per column: one byte `has null buffer` (varint 0/1), ceil(m/8) validity bytes when it
// is 1, then the values: bit-packed for Boolean, raw for 1-byte types, otherwise `width` byte PLANES of m bytes each.
struct ShufCol {
  const void* values;                       // advanced by the Arrow offset (byte-addressable types)
  const uint8_t* validity;                  // may be null for a nullable column that carries no bitmap: every bit is written as 1
  uint32_t bit_offset;                      // Arrow offset for validity / Boolean values
  uint8_t width;                            // 0: Boolean (bits), else 1, 2, 4, 8, 16 bytes
  uint8_t nullable;                         // the field is nullable: the record carries a validity bitmap for it
  uint8_t tma;                              // set by the launcher: the column's tiles are staged by cp.async.bulk (16-byte aligned, width 1/2/4/8)
  uint8_t _pad[1];
  uint32_t k8, kw;                          // columns before this one take k8 * ceil(m/8) + kw * m bytes (+ one flag byte each)
};

struct ShufSpec {
  int32_t ncols, num_partitions, batch_size, nkeys;
  uint32_t tot_k8, tot_kw;                  // record of m rows = varint_len(m) + ncols + tot_k8 * ceil(m/8) + tot_kw * m bytes
  int8_t key_col[8];                        // hash partitioning: indices (into col[]) of the key columns, in hash order
  uint8_t key_phys[8];
  ShufCol col[SHUF_MAX_COLS];
};

__host__ __device__ inline uint32_t shuf_varint_len(unsigned long long m) { uint32_t l = 1; while (m >= 128) { m >>= 7; l++; } return l; }
__host__ __device__ inline unsigned long long shuf_record_bytes(const ShufSpec& sp, unsigned long long m) {
  return shuf_varint_len(m) + (unsigned long long)sp.ncols + (unsigned long long)sp.tot_k8 * ((m + 7) >> 3) + (unsigned long long)sp.tot_kw * m;
}
// bytes of partition holding t rows: records of batch_size rows, the last one shorter
__host__ __device__ inline unsigned long long shuf_partition_bytes(const ShufSpec& sp, unsigned long long t) {
  if (t == 0) return 0;
  const unsigned long long B = (unsigned long long)sp.batch_size, nrec = (t + B - 1) / B;
  return (nrec - 1) * shuf_record_bytes(sp, B) + shuf_record_bytes(sp, t - (nrec - 1) * B);
}

// pids[i] = pmod(murmur3(key columns of row i, seed 42), P) as u16 and counts[p] += 1 (shuffle/mod.rs:163-188); nkeys == 0: every row -> partition 0
int launch_shuffle_pids(const ShufSpec& sp, int64_t n, uint16_t* d_pids, unsigned long long* d_counts /* P, zeroed */, cudaStream_t s);
// part_off[p] = byte offset of partition p in the encoded buffer (P + 1 entries), cursors[p] = 0, record headers + flag bytes written
int launch_shuffle_layout(const ShufSpec& sp, const unsigned long long* d_counts, unsigned long long* d_part_off, unsigned long long* d_cursors, uint8_t* d_out, cudaStream_t s);
// scatter + encode: every value lands in its byte planes
int launch_shuffle_encode(const ShufSpec& sp, const uint16_t* d_pids, int64_t n, const unsigned long long* d_counts, const unsigned long long* d_part_off,
                          unsigned long long* d_cursors, uint8_t* d_out, cudaStream_t s);

}  // namespace b200q
