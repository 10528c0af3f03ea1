// Descriptors of the specialised HashAgg update kernels (kernels_fast.cu).
#pragma once
#include "kernels.cuh"

namespace b200q {

enum FastAccKind : uint8_t { FAST_ACC_ADD = 0, FAST_ACC_COUNT = 1 };   // COUNT with col < 0 is COUNT(*)

struct FastSpec {
  int32_t nkeys, nacc, nfilt, dense;
  int32_t row_kernels;                                    // debugging / A-B measurements: keep the round-1 one-row-per-lane kernels (B200Q_ROW_KERNELS=1)
  int32_t lean, hot_cache;                                // lean: all referenced columns are aligned non-null int64 (set per launch); hot_cache: skewed keys (experimental)
  int8_t key_col[2]; uint8_t key_phys[2];                 // program column slots / physical kinds of the key columns
  struct { uint8_t kind; int8_t col; uint8_t phys; uint8_t vbit; uint8_t word; uint8_t _pad[3]; } acc[2];
  struct { int8_t col; uint8_t phys; uint8_t op; uint8_t _pad[5]; long long lit; } filt[4];
  // the same conjuncts merged per column into closed intervals (tile kernels): row passes iff (u64)(x - lo) <= span for every column
  int32_t nfcol;                                          // -1: the conjuncts cannot be merged (a `!=` term or more than 2 columns): tile kernels not used
  int32_t filt_never;                                     // the merged intervals are empty: no row passes
  struct { int8_t col; uint8_t phys; uint8_t _pad[6]; long long lo; unsigned long long span; } frange[2];
  long long dense_base;                                   // DENSE: entry index = key0 - dense_base                      (one key)
  unsigned long long dense_cap;                           // entries of dense_stride words
  long long dense_base1;                                  //        entry index = (key0 - dense_base) * dense_r1 + (key1 - dense_base1)   (two keys)
  unsigned long long dense_r1, dense_cap0;                // key1 - dense_base1 < dense_r1, key0 - dense_base < dense_cap0; dense_cap = dense_cap0 * dense_r1
  unsigned long long* dense_tab;
  int8_t dense_stride;                                    // 2 or 4 words per entry (= lanes that update one entry in one instruction)
  int8_t dense_word_src[4];                               // per entry word: -1 row counter (+1), -2 padding (+0), j accumulator j, 2+j valid arguments of accumulator j
  uint8_t dense_presence_word;                            // word that is non-zero iff the entry holds a group
  uint8_t _pad1[2];
  unsigned long long* sink;                               // FAST_SINK_WARPS x 4 words: per-warp scratch sector for no-op REDs
};
constexpr int FAST_SINK_WARPS = 4096;

// per emit column: which dense word holds the value and which (count) word validates it
struct DenseEmitMap { uint8_t word[EMIT_MAX_COLS]; uint8_t valid_word[EMIT_MAX_COLS]; };

// dense entry index of a row; false: outside the dense range (the row goes to a hashed slot)
template <int NK>
__device__ __forceinline__ bool dense_index_of(const FastSpec& fs, long long k0, long long k1, unsigned long long& idx) {
  const unsigned long long d0 = (unsigned long long)(k0 - fs.dense_base);
  if (NK == 1) { idx = d0; return d0 < fs.dense_cap; }
  const unsigned long long d1 = (unsigned long long)(k1 - fs.dense_base1);
  idx = d0 * fs.dense_r1 + d1;
  return d0 < fs.dense_cap0 && d1 < fs.dense_r1;
}

int launch_agg_tile_dense(const ColTable& cols, const FastSpec& fs, const AggLayout& lay, const AggTable& tab, int64_t row_begin, int64_t n, cudaStream_t s);
int launch_agg_fast_update(const ColTable& cols, const FastSpec& fs, const AggLayout& lay, const AggTable& tab, int64_t row_begin, int64_t n, cudaStream_t s);
int launch_key_skew_probe(const DevCol* key_cols, const uint8_t* phys, int nkeys, int64_t n, unsigned* d_hist /* 65536 + 1 words, zeroed */, cudaStream_t s);
int launch_key_range(const DevCol& col, int phys, int64_t n, long long* d_out, cudaStream_t s);
int launch_agg_emit_dense(const FastSpec& fs, const EmitTable& emit, const DenseEmitMap& map, unsigned long long* d_out_count, cudaStream_t s);
int launch_dense_count(const FastSpec& fs, unsigned long long* d_out, cudaStream_t s);

}  // namespace b200q
