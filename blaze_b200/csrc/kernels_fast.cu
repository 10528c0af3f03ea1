// Specialised streaming HashAgg update kernels for the shapes TPC-DS q1/q3/q6 produce:
//   * 1-2 integer grouping columns read straight from the Arrow buffers (no row encoding: K4 is gone),
//   * up to 4 fused `column <cmp> literal` conjuncts (the FilterExec below the agg),
//   * 1-2 accumulators of the "64-bit add" class: SUM(int column), COUNT(column), COUNT(*).
//
// What bounds them (profiles/r01_microbench_atomics_*.txt, measured on B200): the input stream runs at
// HBM speed (6.5 TB/s) but every row also needs a random read-modify-write into the L2-resident group
// table, and the chip retires ~1.55e11 scattered 32-byte sector operations per second.  So the design
// goal is ONE sector operation per row:
//   - slot = one 32-byte sector {hdr, key, acc0, acc1}: the probe is a single 16-byte load;
//   - the two accumulators of a row are updated by the SAME `red.add.u64` instruction from two adjacent
//     lanes (lane pairing), which the memory system coalesces into one sector operation;
//   - DENSE mode (single integer key whose values span a small range, e.g. TPC-DS surrogate keys):
//     the slot index is key - base, no probe at all: {rows, acc0, acc1} updated by a gang of 4 lanes in
//     one instruction.  Keys outside the range (and NULL keys) take the hash path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "agg_device.cuh"
#include "kernels_fast.cuh"

namespace b200q {

constexpr int FA_BLOCK = 256;
constexpr int FA_R = 4;                       // rows per thread per tile (independent loads in flight)
constexpr int FA_TILE = FA_BLOCK * FA_R;

// streaming load: bypass L1 and mark the line evict-first in L2 so the input stream does not push the
// group table out of L2
__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); return pol;
}
__device__ __forceinline__ long long ld_stream_s64(const long long* p) {
  long long v; const uint64_t pol = make_evict_first_policy();
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ bool col_valid(const DevCol& c, long long i) {
  if (!c.validity) return true;
  const unsigned long long bi = (unsigned long long)i + c.bit_offset;
  return (__ldg(c.validity + (bi >> 3)) >> (bi & 7)) & 1;
}
__device__ __forceinline__ long long col_load_int(const DevCol& c, int phys, long long i) {
  switch (phys) {
    case PH_I64: return ld_stream_s64((const long long*)c.values + i);
    case PH_I32: return (long long)__ldg((const int32_t*)c.values + i);
    case PH_I16: return (long long)__ldg((const int16_t*)c.values + i);
    case PH_I8: return (long long)__ldg((const int8_t*)c.values + i);
    default: { const unsigned long long bi = (unsigned long long)i + c.bit_offset; return (__ldg((const uint8_t*)c.values + (bi >> 3)) >> (bi & 7)) & 1; }
  }
}
__device__ __forceinline__ bool cmp_apply(int op, long long a, long long b) {
  switch (op) { case CMP_EQ: return a == b; case CMP_NE: return a != b; case CMP_LT: return a < b; case CMP_LE: return a <= b; case CMP_GT: return a > b; default: return a >= b; }
}

template <int NACC, bool DENSE>
__global__ void __launch_bounds__(FA_BLOCK) agg_fast_update_kernel(const ColTable cols, const FastSpec fs, const AggLayout lay, const AggTable tab,
                                                                   long long row_begin, long long n, const uint32_t* __restrict__ row_list) {
  const unsigned lane = threadIdx.x & 31;
  const long long ntiles = (n + FA_TILE - 1) / FA_TILE;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    bool alive[FA_R]; uint32_t rel[FA_R]; long long row[FA_R];
    long long key[2][FA_R]; uint32_t knull[FA_R];
    unsigned long long aval[2][FA_R]; bool aact[2][FA_R];
    // ---- stream the columns: all loads of the tile are issued before any dependent work
#pragma unroll
    for (int r = 0; r < FA_R; r++) {
      const long long i = tile * FA_TILE + r * FA_BLOCK + threadIdx.x;
      alive[r] = i < n;
      rel[r] = alive[r] ? (row_list ? row_list[i] : (uint32_t)i) : 0;
      row[r] = row_begin + rel[r];
    }
#pragma unroll
    for (int r = 0; r < FA_R; r++) {
      knull[r] = 0;
      for (int k = 0; k < fs.nkeys; k++) {
        key[k][r] = 0;
        if (alive[r]) {
          const DevCol& c = cols.col[fs.key_col[k]];
          if (col_valid(c, row[r])) key[k][r] = col_load_int(c, fs.key_phys[k], row[r]); else knull[r] |= 1u << k;
        }
      }
#pragma unroll
      for (int j = 0; j < NACC; j++) {
        aval[j][r] = 1; aact[j][r] = alive[r];
        if (alive[r] && fs.acc[j].col >= 0) {
          const DevCol& c = cols.col[fs.acc[j].col];
          const bool v = col_valid(c, row[r]);
          aact[j][r] = v;
          if (fs.acc[j].kind == FAST_ACC_ADD) aval[j][r] = v ? (unsigned long long)col_load_int(c, fs.acc[j].phys, row[r]) : 0ULL;
        }
      }
    }
    // ---- fused FilterExec conjuncts (null -> false, cached_exprs_evaluator.rs:518-520)
    for (int f = 0; f < fs.nfilt; f++) {
      const DevCol& c = cols.col[fs.filt[f].col];
#pragma unroll
      for (int r = 0; r < FA_R; r++)
        if (alive[r]) alive[r] = col_valid(c, row[r]) && cmp_apply(fs.filt[f].op, col_load_int(c, fs.filt[f].phys, row[r]), fs.filt[f].lit);
    }
    // ---- locate the slot of every row
    unsigned long long* slot[FA_R]; unsigned flags[FA_R]; bool dense[FA_R];
    unsigned long long* cand[FA_R]; ulonglong2 hk[FA_R]; uint64_t h[FA_R];
#pragma unroll
    for (int r = 0; r < FA_R; r++) {
      slot[r] = nullptr; flags[r] = 0; dense[r] = false; cand[r] = nullptr;
      if (!alive[r]) continue;
      if (DENSE) {
        const unsigned long long idx = (unsigned long long)(key[0][r] - fs.dense_base);
        if (knull[r] == 0 && idx < fs.dense_cap) { dense[r] = true; slot[r] = fs.dense_tab + idx * 4; continue; }
      }
      uint64_t kw[2] = {(uint64_t)key[0][r], (uint64_t)key[1][r]};
      h[r] = agg_hash_words(kw, fs.nkeys, knull[r]);
      cand[r] = tab.slots + (h[r] & tab.mask) * (uint64_t)lay.slot_words;
      hk[r] = ld_relaxed_v2u64(cand[r]);                                         // {hdr, key0}: one 16-byte probe
    }
#pragma unroll
    for (int r = 0; r < FA_R; r++) {
      if (!alive[r] || dense[r]) continue;
      const unsigned tag = (unsigned)(h[r] >> 32) | 0x80000000u;
      const unsigned fl = (unsigned)(hk[r].x >> 32);
      bool hit = (unsigned)hk[r].x == tag && (fl >> 16) == knull[r] && hk[r].y == (uint64_t)key[0][r];
      if (hit && fs.nkeys == 2) hit = ld_relaxed_u64(cand[r] + 2) == (uint64_t)key[1][r];
      if (hit) { slot[r] = cand[r]; flags[r] = fl; continue; }
      uint64_t kw[2] = {(uint64_t)key[0][r], (uint64_t)key[1][r]};               // first probe missed: full protocol (insert / walk)
      slot[r] = agg_find_or_insert(lay, tab, kw, knull[r], h[r], &flags[r]);
      if (!slot[r]) { const unsigned long long at = atomicAdd(tab.counters + 1, 1ULL); tab.deferred[at] = rel[r]; alive[r] = false; }
    }
    // ---- accumulate: every lane reaches this point (warp-converged) so lanes can update each other's rows
#pragma unroll
    for (int r = 0; r < FA_R; r++) {
      const bool live = alive[r] && slot[r] != nullptr;
      if (DENSE) {
        // gang of 4 lanes: in step s the 4 lanes update words {rows, acc0, acc1} of the row owned by gang lane s
        const unsigned m = lane & 3, gbase = lane & ~3u;
#pragma unroll
        for (int s = 0; s < 4; s++) {
          const unsigned src = gbase + s;
          const unsigned long long ps = __shfl_sync(0xffffffffu, (unsigned long long)slot[r], src);
          const bool pd = __shfl_sync(0xffffffffu, (int)(live && dense[r]), src);
          const unsigned long long v0 = __shfl_sync(0xffffffffu, aval[0][r], src);
          const bool a0 = __shfl_sync(0xffffffffu, (int)aact[0][r], src);
          unsigned long long v1 = 1; bool a1 = false;
          if (NACC == 2) { v1 = __shfl_sync(0xffffffffu, aval[1][r], src); a1 = __shfl_sync(0xffffffffu, (int)aact[1][r], src); }
          // one predicated RED instruction for the whole gang: the 3 words of a row share a 32-byte sector
          const unsigned long long val = m == 0 ? 1ULL : (m == 1 ? v0 : v1);
          const bool pred = pd && (m == 0 || (m == 1 && a0) || (m == 2 && NACC == 2 && a1));
          if (pred) red_add_u64((unsigned long long*)ps + m, val);
        }
      }
      const bool hashed = live && !dense[r];
      if (NACC == 2) {
        // lane pairing: the owner updates acc0, its neighbour updates acc1 of the same row in the same instruction
        const int w0 = fs.acc[0].word, w1 = fs.acc[1].word;
        const unsigned long long ps = __shfl_xor_sync(0xffffffffu, (unsigned long long)slot[r], 1);
        const bool ph = __shfl_xor_sync(0xffffffffu, (int)hashed, 1);
        const unsigned long long pv1 = __shfl_xor_sync(0xffffffffu, aval[1][r], 1);
        const bool pa1 = __shfl_xor_sync(0xffffffffu, (int)aact[1][r], 1);
        const bool odd = lane & 1;
        {   // step 1: rows owned by even lanes (one predicated RED: both words sit in the same 32-byte sector)
          unsigned long long* ptr = odd ? (unsigned long long*)ps + w1 : slot[r] + w0;
          const unsigned long long val = odd ? pv1 : aval[0][r];
          const bool pred = odd ? (ph && pa1) : (hashed && aact[0][r]);
          if (pred) red_add_u64(ptr, val);
        }
        {   // step 2: rows owned by odd lanes
          unsigned long long* ptr = odd ? slot[r] + w0 : (unsigned long long*)ps + w1;
          const unsigned long long val = odd ? aval[0][r] : pv1;
          const bool pred = odd ? (hashed && aact[0][r]) : (ph && pa1);
          if (pred) red_add_u64(ptr, val);
        }
        if (hashed) { if (aact[0][r]) slot_mark(slot[r], flags[r], fs.acc[0].vbit); if (aact[1][r]) slot_mark(slot[r], flags[r], fs.acc[1].vbit); }
      } else {
        if (hashed && aact[0][r]) { red_add_u64(slot[r] + fs.acc[0].word, aval[0][r]); slot_mark(slot[r], flags[r], fs.acc[0].vbit); }
      }
    }
  }
}

static int fast_grid(int64_t ntiles) {
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t cap = (int64_t)sms * 8;          // persistent grid: a multiple of the SM count
  return (int)(ntiles < cap ? (ntiles < 1 ? 1 : ntiles) : cap);
}

int launch_agg_fast_update(const ColTable& cols, const FastSpec& fs, const AggLayout& lay, const AggTable& tab, int64_t row_begin, int64_t n,
                           const uint32_t* d_row_list, cudaStream_t s) {
  if (n <= 0) return 0;
  const int64_t ntiles = (n + FA_TILE - 1) / FA_TILE;
  const int grid = fast_grid(ntiles);
  const bool dense = fs.dense != 0;
  if (fs.nacc == 2) { if (dense) agg_fast_update_kernel<2, true><<<grid, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n, d_row_list);
                      else agg_fast_update_kernel<2, false><<<grid, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n, d_row_list); }
  else { if (dense) agg_fast_update_kernel<1, true><<<grid, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n, d_row_list);
         else agg_fast_update_kernel<1, false><<<grid, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n, d_row_list); }
  return 1;
}

// ---------------------------------------------------------------------------------------------------
// key range of the first batch (decides DENSE mode)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) key_range_kernel(const DevCol col, int phys, long long n, long long* out /*[0]=min,[1]=max,[2]=non-null count*/) {
  long long mn = INT64_MAX, mx = INT64_MIN; unsigned long long cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!col_valid(col, i)) continue;
    const long long v = col_load_int(col, phys, i);
    mn = v < mn ? v : mn; mx = v > mx ? v : mx; cnt++;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
    mn = a < mn ? a : mn; mx = b > mx ? b : mx; cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(out, mn); atomicMax(out + 1, mx); atomicAdd((unsigned long long*)out + 2, cnt); }
}
int launch_key_range(const DevCol& col, int phys, int64_t n, long long* d_out, cudaStream_t s) {
  key_range_kernel<<<fast_grid((n + 2047) / 2048), 256, 0, s>>>(col, phys, n, d_out);
  return 1;
}

// ---------------------------------------------------------------------------------------------------
// emit of the dense table (same output columns as agg_emit_kernel)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dense_store(const EmitCol& c, unsigned long long at, uint64_t lo, bool valid) {
  if (c.valid_bytes) c.valid_bytes[at] = valid ? 1 : 0;
  switch (c.phys) {
    case PH_BOOL: ((uint8_t*)c.values)[at] = lo != 0; break;
    case PH_I8: ((int8_t*)c.values)[at] = (int8_t)lo; break;
    case PH_I16: ((int16_t*)c.values)[at] = (int16_t)lo; break;
    case PH_I32: ((int32_t*)c.values)[at] = (int32_t)lo; break;
    default: ((uint64_t*)c.values)[at] = lo; break;
  }
}
__global__ void __launch_bounds__(256) agg_emit_dense_kernel(const FastSpec fs, const EmitTable emit, const DenseEmitMap map, unsigned long long* out_count) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounds = (fs.dense_cap + stride - 1) / stride;
  for (uint64_t it = 0; it < rounds; it++) {
    const uint64_t i = it * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const unsigned long long* e = fs.dense_tab + i * 4;
    const bool occ = i < fs.dense_cap && e[0] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    if (!m) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(out_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!occ) continue;
    const unsigned long long at = base + __popc(m & lanemask_lt());
    for (int c = 0; c < emit.ncols; c++) {
      const EmitCol ec = emit.col[c];
      if (ec.kind == EMIT_KEY) dense_store(ec, at, (uint64_t)(fs.dense_base + (long long)i), true);
      else {
        const int w = map.word[c], vw = map.valid_word[c];
        const bool valid = vw == 0xFF ? true : e[vw] != 0;
        dense_store(ec, at, valid ? e[w] : 0, valid);
      }
    }
  }
}
int launch_agg_emit_dense(const FastSpec& fs, const EmitTable& emit, const DenseEmitMap& map, unsigned long long* d_out_count, cudaStream_t s) {
  agg_emit_dense_kernel<<<fast_grid(((int64_t)fs.dense_cap + 255) / 256), 256, 0, s>>>(fs, emit, map, d_out_count);
  return 1;
}

// number of occupied dense entries (table growth accounting is not needed: the dense table never fills)
__global__ void __launch_bounds__(256) dense_count_kernel(const unsigned long long* tab, uint64_t cap, unsigned long long* out) {
  unsigned long long c = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) c += tab[i * 4] != 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
int launch_dense_count(const unsigned long long* tab, uint64_t cap, unsigned long long* d_out, cudaStream_t s) {
  dense_count_kernel<<<fast_grid(((int64_t)cap + 255) / 256), 256, 0, s>>>(tab, cap, d_out);
  return 1;
}

}  // namespace b200q
