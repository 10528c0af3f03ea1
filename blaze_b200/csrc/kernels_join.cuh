// Descriptors and launchers of the hash-join kernels (kernels_join.cu): build, probe (count + emit), gather.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace b200q {

constexpr uint32_t JOIN_NIL = 0xFFFFFFFFu;

// key columns of one side: up to two integer-like columns (int8 .. int64, date32, timestamp), each widened to i64
struct JoinKeys {
  int32_t nkeys;
  uint8_t phys[2];
  uint8_t _pad[2];
  DevCol col[2];
};

// Build side table (replaces JoinHashMap, joins/join_hash_map.rs:91-275): one slot per DISTINCT key holding the head of
// a chain through `next[]` of the build rows with that key and their number; rows with a NULL key are not inserted
// (:119-128).  Open addressing, linear probing, capacity a power of two >= 2 x rows (load <= 0.5 like the reference's
// `max(128) * 2 / 8` groups, :147-150).
struct JoinTable {
  unsigned long long* keys;     // capacity x nkw words
  uint32_t* state;              // 0 empty, 1 being written, 2 ready
  uint32_t* head;               // first build row of the key's chain
  uint32_t* count;              // rows in the chain
  uint32_t* next;               // per build row: next row with the same key, JOIN_NIL at the end
  uint32_t* stats;              // [0] rows of the most duplicated key (0 or 1: the keys are unique)
  unsigned long long* packed;   // probe view, written once after the build: per slot {key words..., count << 32 | head} (2 or 4 words; head = NIL: empty slot)
  uint32_t mask;                // capacity - 1
  int32_t nkw;                  // key words per slot (1 or 2)
};

int launch_join_build(const JoinKeys& k, int64_t n, const JoinTable& t, cudaStream_t s);      // fills keys / state / head / count / next, then packs the probe view
// per probe row: head of the matching chain (JOIN_NIL: no match / NULL key) and out_count = matches (probe_outer: at least 1)
int launch_join_probe_count(const JoinKeys& k, int64_t n, const JoinTable& t, int probe_outer, uint32_t* d_head, int32_t* d_count, cudaStream_t s, unsigned long long* d_total = nullptr);   // d_total += sum of the counts
// (probe row, build row) pairs at offs[r]...; marks map_joined[build row] = 1 when mark != null; unmatched outer rows pair with JOIN_NIL
int launch_join_probe_emit(int64_t n, const JoinTable& t, const uint32_t* d_head, const int32_t* d_offs, uint32_t* d_pidx, uint32_t* d_bidx, uint8_t* mark, cudaStream_t s);
// fused probes (no per-row intermediates; output positions reserved with one atomic per warp on *d_cursor, zeroed by the caller):
//   pairs : (probe row, build row) pairs of Inner / Left / Right / Full; d_pidx == null only counts
//   select: the probe rows with (invert: without) a partner — LeftSemi / LeftAnti probed from the left
//   mark  : marks the build rows whose key some probe row has — semi forms where the build side is the join side
int launch_join_probe_pairs(const JoinKeys& k, int64_t n, const JoinTable& t, int probe_outer, unsigned long long* d_cursor, uint32_t* d_pidx, uint32_t* d_bidx, uint8_t* mark, cudaStream_t s);
int launch_join_probe_select(const JoinKeys& k, int64_t n, const JoinTable& t, int invert, unsigned long long* d_cursor, uint32_t* d_idx, cudaStream_t s);
int launch_join_probe_mark(const JoinKeys& k, int64_t n, const JoinTable& t, uint8_t* mark, cudaStream_t s);
// semi-style probes where the BUILD side is the join side: mark every build row whose key some probe row has
int launch_join_mark_build(int64_t n, const JoinTable& t, const uint32_t* d_head, uint8_t* mark, cudaStream_t s);
// out[i] = idx[i] == JOIN_NIL ? NULL : src[idx[i]]   (width bytes per value; src_valid / out_valid: one byte per row, may be null)
int launch_join_gather(const void* src, const uint8_t* src_valid_bits, uint32_t src_bit_offset, const uint8_t* src_valid_bytes, int width, const uint32_t* idx, int64_t n,
                       void* out, uint8_t* out_valid_bytes, cudaStream_t s);
// the same for up to 16 columns that share one index vector, in one pass
struct GatherCol { const void* src; const uint8_t* vbits; const uint8_t* vbytes; void* out; uint8_t* out_valid; uint32_t bit_offset; int32_t width; };
struct GatherSpec { int32_t ncols; int32_t _pad; GatherCol col[16]; };
int launch_join_gather_multi(const GatherSpec& g, const uint32_t* idx, int64_t n, cudaStream_t s);
// unique map keys: gather of both sides fused, driven by the per-row chain heads of launch_join_probe_count (d_head), outputs written in probe-row order at
// positions reserved per tile on *d_cursor; probe_cols: src = the probe
// batch's columns (vbits + bit_offset for validity), build_cols: src = the map side's columns (vbytes); at most 16 columns per side
int launch_join_probe_fused(const uint32_t* d_head, int64_t n, int probe_outer, unsigned long long* d_cursor, const GatherSpec& probe_cols, const GatherSpec& build_cols,
                            uint8_t* mark, cudaStream_t s);
// bytes[i] = bit (i + bit_offset) of bits (all 1 when bits is null)
int launch_unpack_bits(const uint8_t* bits, uint32_t bit_offset, int64_t n, uint8_t* bytes, cudaStream_t s);
// flags[i] = (head[i] != NIL) ^ invert, as int32 for the scan; idx[offs[i]] = i for rows whose flag is set
int launch_join_flags(const uint32_t* d_head, int64_t n, int invert, int32_t* d_flags, cudaStream_t s);
int launch_join_compact_indices(const int32_t* d_flags, const int32_t* d_offs, int64_t n, uint32_t* d_idx, cudaStream_t s);
int launch_join_match_bytes(const uint32_t* d_head, int64_t n, uint8_t* d_bytes, cudaStream_t s);      // bytes[i] = row i found a match
int launch_bytes_to_flags(const uint8_t* bytes, int64_t n, int invert, int32_t* d_flags, cudaStream_t s);

}  // namespace b200q
