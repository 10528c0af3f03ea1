// Specialised streaming HashAgg update kernels for the shapes TPC-DS q1/q3/q6 produce:
//   * 1-2 integer grouping columns read straight from the Arrow buffers (no row encoding: K4 is gone),
//   * up to 4 fused `column <cmp> literal` conjuncts (the FilterExec below the agg),
//   * 1-2 accumulators of the "64-bit add" class: SUM(int column), COUNT(column), COUNT(*).
//
// What bounds them (profiles/r01_microbench_atomics_*.txt, measured on B200): the input stream runs at
// HBM speed (6.5 TB/s) but every row also needs a random read-modify-write into the L2-resident group
// table, and the chip retires ~1.55e11 scattered 32-byte sector operations per second.  So the design
// goal is ONE sector operation per row:
//   - hashed: slot = one 32-byte sector {hdr, key, acc0, acc1}: the probe is a single 16-byte load and the
//     two accumulators of a row are updated by the SAME red.add.u64 instruction from two adjacent lanes,
//     which the memory system coalesces into one sector operation;
//   - DENSE (single integer key whose values span a small range, e.g. TPC-DS surrogate keys): the entry
//     index is key - base, no probe at all; an entry is 2 or 4 words ({sum,count} / {rows,acc0,acc1,-})
//     updated by a gang of 2 or 4 lanes in one instruction.  Keys outside the range and NULL keys take
//     the hash table.
//
// Lane organisation ("gang"): G adjacent lanes own G consecutive rows; EVERY lane of the gang loads the
// gang's G rows itself (the lanes read the same addresses: one access for the LSU), computes the G entry
// addresses redundantly, and in step s = 0..G-1 the gang's lanes update word 0..G-1 of row s.  No shuffles,
// no per-thread arrays with dynamic indices.
//
// REDs are issued UNCONDITIONALLY: ptxas if-converts a predicated `red` into `@P ATOMG ... RZ` (an atomic
// with a return path); lanes with nothing to add send +0 to their warp's private sink sector instead.
#include <cuda_runtime.h>
#include <stdint.h>

#include "agg_device.cuh"
#include "kernels_fast.cuh"

namespace b200q {

constexpr int FA_BLOCK = 256;
constexpr int FA_UNITS = 4;                   // 32-row units per warp per tile (generic gang kernel)
constexpr int FA_TILE = FA_BLOCK * FA_UNITS;

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
__device__ __forceinline__ unsigned long long* warp_sink(const FastSpec& fs, long long gwarp, unsigned m) {
  return fs.sink + ((gwarp & (FAST_SINK_WARPS - 1)) << 2) + (m & 3);
}

// ---------------------------------------------------------------------------------------------------
// generic gang kernel: any integer widths, validity bitmaps, 1-2 keys; DG = dense gang width (0: no dense table)
// ---------------------------------------------------------------------------------------------------
template <int NK, int NACC, int DG>
__global__ void __launch_bounds__(FA_BLOCK) agg_gang_update_kernel(const ColTable cols, const FastSpec fs, const AggLayout lay, const AggTable tab,
                                                                   long long row_begin, long long n) {
  constexpr bool DENSE = DG != 0;
  constexpr int G = DENSE ? DG : (NACC == 2 ? 2 : 1);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = lane % G, gl = lane - m;                   // my word index inside the gang, first lane of my gang
  const long long ntiles = (n + FA_TILE - 1) / FA_TILE;
  // which accumulator (if any) this lane updates: dense entries follow fs.dense_word_src, hashed slots acc m
  const int src = DENSE ? fs.dense_word_src[m] : (int)m;        // -1: row counter (+1), -2: padding (+0), j: accumulator j
  const bool has_acc = src >= 0 && src < NACC;
  const int acc_col = has_acc ? fs.acc[src].col : -1;
  const int acc_kind = has_acc ? fs.acc[src].kind : FAST_ACC_COUNT;
  const int acc_phys = has_acc ? fs.acc[src].phys : PH_I64;
  const int acc_word = has_acc ? fs.acc[src].word : 0;
  const int acc_vbit = has_acc ? fs.acc[src].vbit : 0xFF;
  unsigned long long* const sink = warp_sink(fs, (long long)blockIdx.x * (FA_BLOCK / 32) + warp, m);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
    for (int u = 0; u < FA_UNITS; u++) {
      const long long rel0 = tile * FA_TILE + (long long)(u * (FA_BLOCK / 32) + warp) * 32 + gl;   // first row of my gang (relative)
      long long key0[G], key1[G]; bool alive[G]; unsigned knull[G];
      unsigned long long val[G]; bool act[G];
#pragma unroll
      for (int s = 0; s < G; s++) {
        const long long rel = rel0 + s, row = row_begin + rel;
        alive[s] = rel < n; knull[s] = 0; key0[s] = 0; key1[s] = 0; val[s] = src == -2 ? 0 : 1; act[s] = alive[s];
        if (!alive[s]) continue;
        { const DevCol& c = cols.col[fs.key_col[0]]; if (col_valid(c, row)) key0[s] = col_load_int(c, fs.key_phys[0], row); else knull[s] |= 1u; }
        if (NK == 2) { const DevCol& c = cols.col[fs.key_col[1]]; if (col_valid(c, row)) key1[s] = col_load_int(c, fs.key_phys[1], row); else knull[s] |= 2u; }
        if (has_acc && acc_col >= 0) {
          const DevCol& c = cols.col[acc_col];
          act[s] = col_valid(c, row);
          if (acc_kind == FAST_ACC_ADD) val[s] = act[s] ? (unsigned long long)col_load_int(c, acc_phys, row) : 0ULL;
        }
      }
      // fused FilterExec conjuncts (null -> false, cached_exprs_evaluator.rs:518-520)
      for (int f = 0; f < fs.nfilt; f++) {
        const DevCol& c = cols.col[fs.filt[f].col];
#pragma unroll
        for (int s = 0; s < G; s++) {
          const long long row = row_begin + rel0 + s;
          if (alive[s]) alive[s] = col_valid(c, row) && cmp_apply(fs.filt[f].op, col_load_int(c, fs.filt[f].phys, row), fs.filt[f].lit);
        }
      }
      // entry / slot of every row of the gang (computed redundantly by each lane of the gang)
      unsigned long long* ptr[G]; unsigned long long* slot[G] /* key entry of a hashed row */; unsigned flags[G]; bool need[G], ins[G]; uint64_t h[G], idx[G];
#pragma unroll
      for (int s = 0; s < G; s++) {
        ptr[s] = nullptr; slot[s] = nullptr; flags[s] = 0; need[s] = false; ins[s] = false; h[s] = 0; idx[s] = 0;
        if (!alive[s]) continue;
        if (DENSE) {
          const unsigned long long di = (unsigned long long)(key0[s] - fs.dense_base);
          if (knull[s] == 0 && di < fs.dense_cap) { ptr[s] = fs.dense_tab + di * G + m; continue; }
        }
        h[s] = agg_hash2((uint64_t)key0[s], NK == 2 ? (uint64_t)key1[s] : 0ULL, knull[s]);
        idx[s] = agg_first_slot(h[s], tab.capacity); need[s] = true;
      }
      // probe walk: all pending rows advance one slot per round, the lanes of a gang in lockstep (collisions are
      // common at load 0.5, so they must not serialise the warp); only NEW keys go to the insert section
      while (true) {
        ulonglong2 hk[G];
#pragma unroll
        for (int s = 0; s < G; s++) if (need[s]) hk[s] = ld_relaxed_v2u64(tab.keys + idx[s] * (uint64_t)lay.kstride);   // {hdr, key0}
        bool pending = false;
#pragma unroll
        for (int s = 0; s < G; s++) {
          if (!need[s]) continue;
          const unsigned tag = agg_tag(h[s]), t = (unsigned)hk[s].x;
          unsigned long long* sp = tab.keys + idx[s] * (uint64_t)lay.kstride;
          if (t == tag) {
            bool hit = (unsigned)(hk[s].x >> 48) == knull[s] && hk[s].y == (uint64_t)key0[s];
            if (NK == 2 && hit) hit = ld_relaxed_u64(sp + 2) == (uint64_t)key1[s];
            if (hit) { slot[s] = sp; flags[s] = (unsigned)(hk[s].x >> 32); need[s] = false; }
            else idx[s] = agg_next_slot(idx[s], tab.capacity);
          } else if (t == TAG_EMPTY) { need[s] = false; ins[s] = true; }
          else if (t != TAG_LOCKED) idx[s] = agg_next_slot(idx[s], tab.capacity);
          pending |= need[s];
        }
        if (!__any_sync(0xffffffffu, pending)) break;
      }
      bool any_ins = false;
#pragma unroll
      for (int s = 0; s < G; s++) any_ins |= ins[s];
      if (__any_sync(0xffffffffu, any_ins)) {                  // new keys: lane 0 of the gang inserts, then broadcasts
#pragma unroll
        for (int s = 0; s < G; s++) {
          unsigned long long si = idx[s]; unsigned fl = flags[s]; bool inserted = false;
          if (ins[s] && m == 0) {
            uint64_t kw[2] = {(uint64_t)key0[s], (uint64_t)key1[s]};
            si = agg_find_or_insert(lay, tab, kw, knull[s], h[s], &fl, &inserted);
            if (si == AGG_NO_SLOT) { const unsigned long long at = atomicAdd(tab.counters + 1, 1ULL); tab.deferred[at] = (uint32_t)(rel0 + s); }
          }
          { const unsigned b = __ballot_sync(0xffffffffu, inserted); if (lane == 0 && b) atomicAdd(tab.counters, (unsigned long long)__popc(b)); }   // one counter update per warp step
          if (G > 1) { si = __shfl_sync(0xffffffffu, si, gl); fl = __shfl_sync(0xffffffffu, fl, gl); }
          if (ins[s]) { idx[s] = si; flags[s] = fl; slot[s] = si == AGG_NO_SLOT ? nullptr : tab.keys + si * (uint64_t)lay.kstride; if (si == AGG_NO_SLOT) alive[s] = false; }
        }
      }
#pragma unroll
      for (int s = 0; s < G; s++)
        if (alive[s] && slot[s] && has_acc) ptr[s] = tab.accs + idx[s] * (uint64_t)lay.astride + acc_word;   // hashed row: this lane's accumulator word
      // accumulate: step s updates row s; the gang's lanes hit adjacent words of ONE sector in ONE instruction
#pragma unroll
      for (int s = 0; s < G; s++) {
        const bool pred = alive[s] && ptr[s] != nullptr;
        red_add_u64(pred ? ptr[s] : sink, (pred && act[s]) ? val[s] : 0ULL);
        if (pred && act[s] && slot[s]) slot_mark(slot[s], flags[s], acc_vbit);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LEAN dense kernel: the hot loop of the M1 shape with everything resolved at compile time.
// Preconditions (checked on the host per launch): one int64 key column, accumulator/filter columns int64,
// none of them carries a validity bitmap, base pointers 32-byte aligned, row_begin % 4 == 0.
// Per 32 rows a warp issues 1 (+1) wide streaming loads and G REDG: ~2 instructions per row.
// ---------------------------------------------------------------------------------------------------
template <int G> struct i64xG { long long v[G]; };
__device__ __forceinline__ i64xG<4> ld_stream_vec(const long long* p, i64xG<4>*) {
  i64xG<4> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ i64xG<2> ld_stream_vec(const long long* p, i64xG<2>*) {
  i64xG<2> r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(r.v[0]), "=l"(r.v[1]) : "l"(p));
  return r;
}
__device__ __forceinline__ i64xG<1> ld_stream_vec(const long long* p, i64xG<1>*) {
  i64xG<1> r;
  asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(r.v[0]) : "l"(p));
  return r;
}
template <int G>
__device__ __forceinline__ i64xG<G> ld_rows(const long long* col, long long rel0, long long n) {
  if (rel0 + G <= n) return ld_stream_vec(col + rel0, (i64xG<G>*)nullptr);
  i64xG<G> r;
#pragma unroll
  for (int s = 0; s < G; s++) r.v[s] = rel0 + s < n ? __ldg(col + rel0 + s) : 0;
  return r;
}

template <int NACC, int G>
__global__ void __launch_bounds__(FA_BLOCK) agg_lean_dense_kernel(const ColTable cols, const FastSpec fs, const AggLayout lay, const AggTable tab,
                                                                  long long row_begin, long long n) {
  constexpr int U = G == 2 ? 4 : 2;                             // 32-row units in flight per warp
  const unsigned lane = threadIdx.x & 31, m = lane % G, gl = lane - m;
  const long long gwarp = (long long)blockIdx.x * (FA_BLOCK / 32) + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * (FA_BLOCK / 32);
  const long long nunits = (n + 31) / 32;
  const long long* kcol = (const long long*)cols.col[fs.key_col[0]].values + row_begin;
  const int src = fs.dense_word_src[m];                         // -1: row counter (+1), -2: padding (+0), j: accumulator j
  const bool has_acc = src >= 0 && src < NACC;
  const bool is_add = has_acc && fs.acc[has_acc ? src : 0].kind == FAST_ACC_ADD;
  const long long* vcol = is_add ? (const long long*)cols.col[fs.acc[src].col].values + row_begin : nullptr;
  const long long cst = src == -2 ? 0 : 1;
  unsigned long long* const sink = warp_sink(fs, gwarp, m);
  const long long base = fs.dense_base; const unsigned long long cap = fs.dense_cap;
  unsigned long long* const dtab = fs.dense_tab + m;

  for (long long unit0 = gwarp * U; unit0 < nunits; unit0 += nwarps * U) {
    i64xG<G> k[U], v[U]; bool alive[U][G];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long rel0 = (unit0 + u) * 32 + gl;
      k[u] = ld_rows<G>(kcol, rel0, n);
      if (is_add) v[u] = ld_rows<G>(vcol, rel0, n);
      else {
#pragma unroll
        for (int s = 0; s < G; s++) v[u].v[s] = cst;
      }
#pragma unroll
      for (int s = 0; s < G; s++) alive[u][s] = rel0 + s < n;
    }
    for (int f = 0; f < fs.nfilt; f++) {                        // fused FilterExec conjuncts
      const long long* fcol = (const long long*)cols.col[fs.filt[f].col].values + row_begin;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const i64xG<G> x = ld_rows<G>(fcol, (unit0 + u) * 32 + gl, n);
#pragma unroll
        for (int s = 0; s < G; s++) alive[u][s] = alive[u][s] && cmp_apply(fs.filt[f].op, x.v[s], fs.filt[f].lit);
      }
    }
    bool oor = false;
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
      for (int s = 0; s < G; s++) {
        const unsigned long long idx = (unsigned long long)(k[u].v[s] - base);
        const bool in = alive[u][s] && idx < cap;
        oor |= alive[u][s] && !in;
        red_add_u64(in ? dtab + idx * G : sink, in ? (unsigned long long)v[u].v[s] : 0ULL);     // G lanes -> 1 sector
      }
    }
    // keys outside the dense range (rare): lane 0 of the gang routes the row through the hash table
    if (__any_sync(0xffffffffu, oor)) {
#pragma unroll
      for (int u = 0; u < U; u++) {
#pragma unroll
        for (int s = 0; s < G; s++) {
          const unsigned long long idx = (unsigned long long)(k[u].v[s] - base);
          bool inserted = false;
          if (alive[u][s] && idx >= cap && m == 0) {
            const long long rel = (unit0 + u) * 32 + gl + s;
            uint64_t kw[2] = {(uint64_t)k[u].v[s], 0};
            unsigned fl;
            const uint64_t si = agg_find_or_insert(lay, tab, kw, 0, agg_hash_words(kw, 1, 0), &fl, &inserted);
            if (si == AGG_NO_SLOT) { const unsigned long long at = atomicAdd(tab.counters + 1, 1ULL); tab.deferred[at] = (uint32_t)rel; }
            else {
              unsigned long long* const p = tab.accs + si * (uint64_t)lay.astride;
#pragma unroll
              for (int j = 0; j < NACC; j++) {
                const unsigned long long x = fs.acc[j].kind == FAST_ACC_ADD ? (unsigned long long)__ldg((const long long*)cols.col[fs.acc[j].col].values + row_begin + rel) : 1ULL;
                atomicAdd(p + fs.acc[j].word, x);
                slot_mark(tab.keys + si * (uint64_t)lay.kstride, fl, fs.acc[j].vbit);
              }
            }
          }
          const unsigned b = __ballot_sync(0xffffffffu, inserted);
          if (lane == 0 && b) atomicAdd(tab.counters, (unsigned long long)__popc(b));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LEAN hashed kernel: 1-2 non-null int64 keys, non-null int64 accumulator/filter columns (same preconditions
// as the lean dense kernel).  One row per lane (the hash + probe work is done once per row); with two
// accumulators the two words of a row are still updated by ONE instruction: neighbouring lanes exchange their
// slot pointer / second operand with one shuffle pair and lane L updates acc0 of its own row while lane L^1
// updates acc1 of the same row (step 1: rows of even lanes, step 2: rows of odd lanes).
// ---------------------------------------------------------------------------------------------------
template <int NK, int NACC>
__global__ void __launch_bounds__(FA_BLOCK) agg_lean_hash_kernel(const ColTable cols, const FastSpec fs, const AggLayout lay, const AggTable tab,
                                                                 long long row_begin, long long n) {
  constexpr int U = 4;                                          // rows per lane in flight
  const unsigned lane = threadIdx.x & 31;
  const bool odd = lane & 1;
  const long long gwarp = (long long)blockIdx.x * (FA_BLOCK / 32) + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * (FA_BLOCK / 32);
  const long long nunits = (n + 31) / 32;
  const long long* kcol0 = (const long long*)cols.col[fs.key_col[0]].values + row_begin;
  const long long* kcol1 = NK == 2 ? (const long long*)cols.col[fs.key_col[1]].values + row_begin : nullptr;
  const long long* vcol0 = fs.acc[0].kind == FAST_ACC_ADD ? (const long long*)cols.col[fs.acc[0].col].values + row_begin : nullptr;
  const long long* vcol1 = (NACC == 2 && fs.acc[1].kind == FAST_ACC_ADD) ? (const long long*)cols.col[fs.acc[1].col].values + row_begin : nullptr;
  const int w0 = fs.acc[0].word, w1 = NACC == 2 ? fs.acc[1].word : 0;
  unsigned long long* const sink = warp_sink(fs, gwarp, lane);
  const uint64_t cap = tab.capacity; const int ks = lay.kstride, as = lay.astride;

  for (long long unit0 = gwarp * U; unit0 < nunits; unit0 += nwarps * U) {
    long long k0[U], k1[U]; unsigned long long v0[U], v1[U]; bool alive[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long rel = (unit0 + u) * 32 + lane;
      alive[u] = rel < n;
      k0[u] = alive[u] ? ld_stream_vec(kcol0 + rel, (i64xG<1>*)nullptr).v[0] : 0;
      k1[u] = (NK == 2 && alive[u]) ? ld_stream_vec(kcol1 + rel, (i64xG<1>*)nullptr).v[0] : 0;
      v0[u] = (vcol0 && alive[u]) ? (unsigned long long)ld_stream_vec(vcol0 + rel, (i64xG<1>*)nullptr).v[0] : 1ULL;
      v1[u] = (vcol1 && alive[u]) ? (unsigned long long)ld_stream_vec(vcol1 + rel, (i64xG<1>*)nullptr).v[0] : 1ULL;
    }
    for (int f = 0; f < fs.nfilt; f++) {                        // fused FilterExec conjuncts
      const long long* fcol = (const long long*)cols.col[fs.filt[f].col].values + row_begin;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long rel = (unit0 + u) * 32 + lane;
        const long long x = rel < n ? ld_stream_vec(fcol + rel, (i64xG<1>*)nullptr).v[0] : 0;
        alive[u] = alive[u] && cmp_apply(fs.filt[f].op, x, fs.filt[f].lit);
      }
    }
    // probe walk: every pending row advances one slot per round; collisions are common at load 0.5 (~25 % of the
    // first probes) so the walk is lane-parallel; only genuinely NEW keys take the insert section
    unsigned long long* slot[U]; uint64_t h[U], idx[U]; unsigned flags[U]; bool need[U], ins[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      h[u] = agg_hash2((uint64_t)k0[u], NK == 2 ? (uint64_t)k1[u] : 0ULL, 0);
      idx[u] = agg_first_slot(h[u], cap); need[u] = alive[u]; ins[u] = false; slot[u] = nullptr; flags[u] = 0;
    }
    while (true) {
      ulonglong2 hk[U];
#pragma unroll
      for (int u = 0; u < U; u++) if (need[u]) hk[u] = ld_relaxed_v2u64(tab.keys + idx[u] * (uint64_t)ks);     // {hdr, key0}: one 16-byte probe
      bool pending = false;
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (!need[u]) continue;
        const unsigned tag = agg_tag(h[u]), t = (unsigned)hk[u].x;
        unsigned long long* sp = tab.keys + idx[u] * (uint64_t)ks;
        if (t == tag) {
          bool hit = (unsigned)(hk[u].x >> 48) == 0 && hk[u].y == (uint64_t)k0[u];
          if (NK == 2 && hit) hit = ld_relaxed_u64(sp + 2) == (uint64_t)k1[u];
          if (hit) { slot[u] = sp; flags[u] = (unsigned)(hk[u].x >> 32); need[u] = false; }
          else idx[u] = agg_next_slot(idx[u], cap);
        } else if (t == TAG_EMPTY) { need[u] = false; ins[u] = true; }
        else if (t != TAG_LOCKED) idx[u] = agg_next_slot(idx[u], cap);      // another key: next slot (locked: look again)
        pending |= need[u];
      }
      if (!__any_sync(0xffffffffu, pending)) break;
    }
    bool any_ins = false;
#pragma unroll
    for (int u = 0; u < U; u++) any_ins |= ins[u];
    if (__any_sync(0xffffffffu, any_ins)) {                     // new keys: full insert protocol, one counter update per warp step
#pragma unroll
      for (int u = 0; u < U; u++) {
        bool inserted = false;
        if (ins[u]) {
          uint64_t kw[2] = {(uint64_t)k0[u], NK == 2 ? (uint64_t)k1[u] : 0ULL};
          idx[u] = agg_find_or_insert(lay, tab, kw, 0, h[u], &flags[u], &inserted);
          if (idx[u] == AGG_NO_SLOT) { const unsigned long long at = atomicAdd(tab.counters + 1, 1ULL); tab.deferred[at] = (uint32_t)((unit0 + u) * 32 + lane); alive[u] = false; slot[u] = nullptr; }
          else slot[u] = tab.keys + idx[u] * (uint64_t)ks;
        }
        const unsigned b = __ballot_sync(0xffffffffu, inserted);
        if (lane == 0 && b) atomicAdd(tab.counters, (unsigned long long)__popc(b));
      }
    }
    // accumulate (REDs unconditional: idle lanes add 0 to the warp's sink sector)
#pragma unroll
    for (int u = 0; u < U; u++) {
      const bool live = alive[u] && slot[u] != nullptr;
      unsigned long long* const ae = tab.accs + idx[u] * (uint64_t)as;          // accumulator entry (never read here: RED only)
      if (NACC == 1) {
        red_add_u64(live ? ae + w0 : sink, live ? v0[u] : 0ULL);
      } else {
        const unsigned long long ps = __shfl_xor_sync(0xffffffffu, live ? (unsigned long long)ae : 0ULL, 1);
        const unsigned long long pv1 = __shfl_xor_sync(0xffffffffu, v1[u], 1);
        unsigned long long* const mine = live ? ae + w0 : sink;  const unsigned long long mv = live ? v0[u] : 0ULL;
        unsigned long long* const theirs = ps ? (unsigned long long*)ps + w1 : sink;  const unsigned long long tv = ps ? pv1 : 0ULL;
        red_add_u64(odd ? theirs : mine, odd ? tv : mv);        // step 1: rows of even lanes: {acc0 by the owner, acc1 by its odd neighbour}
        red_add_u64(odd ? mine : theirs, odd ? mv : tv);        // step 2: rows of odd lanes
      }
      if (live) { slot_mark(slot[u], flags[u], fs.acc[0].vbit); if (NACC == 2) slot_mark(slot[u], flags[u], fs.acc[1].vbit); }
    }
  }
}

static int fast_grid(int64_t ntiles) {
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t cap = (int64_t)sms * 8;          // persistent grid: a multiple of the SM count
  return (int)(ntiles < cap ? (ntiles < 1 ? 1 : ntiles) : cap);
}

template <int NK, int NACC>
static void launch_gang(int dg, int grid, cudaStream_t s, const ColTable& cols, const FastSpec& fs, const AggLayout& lay, const AggTable& tab, int64_t row_begin, int64_t n) {
  if (dg == 4) agg_gang_update_kernel<NK, NACC, 4><<<grid, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n);
  else if (dg == 2) agg_gang_update_kernel<NK, NACC, 2><<<grid, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n);
  else agg_gang_update_kernel<NK, NACC, 0><<<grid, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n);
}

int launch_agg_fast_update(const ColTable& cols, const FastSpec& fs, const AggLayout& lay, const AggTable& tab, int64_t row_begin, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  const int dg = fs.dense ? fs.dense_stride : 0;
  if (dg && fs.lean) {
    const int u = dg == 2 ? 4 : 2;
    const int g = fast_grid((n + 32 * 8 * u - 1) / (32 * 8 * u));
    if (fs.nacc == 2) { if (dg == 2) agg_lean_dense_kernel<2, 2><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); else agg_lean_dense_kernel<2, 4><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); }
    else { if (dg == 2) agg_lean_dense_kernel<1, 2><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); else agg_lean_dense_kernel<1, 4><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); }
    return 1;
  }
  if (!dg && fs.lean) {
    const int g = fast_grid((n + 32 * 8 * 4 - 1) / (32 * 8 * 4));
    if (fs.nkeys == 1) { if (fs.nacc == 2) agg_lean_hash_kernel<1, 2><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); else agg_lean_hash_kernel<1, 1><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); }
    else { if (fs.nacc == 2) agg_lean_hash_kernel<2, 2><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); else agg_lean_hash_kernel<2, 1><<<g, FA_BLOCK, 0, s>>>(cols, fs, lay, tab, row_begin, n); }
    return 1;
  }
  const int grid = fast_grid((n + FA_TILE - 1) / FA_TILE);
  if (fs.nkeys == 1) { if (fs.nacc == 2) launch_gang<1, 2>(dg, grid, s, cols, fs, lay, tab, row_begin, n); else launch_gang<1, 1>(dg, grid, s, cols, fs, lay, tab, row_begin, n); }
  else { if (fs.nacc == 2) launch_gang<2, 2>(0, grid, s, cols, fs, lay, tab, row_begin, n); else launch_gang<2, 1>(0, grid, s, cols, fs, lay, tab, row_begin, n); }
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
    const unsigned long long* e = fs.dense_tab + i * fs.dense_stride;
    const bool occ = i < fs.dense_cap && e[fs.dense_presence_word] != 0;
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

// number of occupied dense entries
__global__ void __launch_bounds__(256) dense_count_kernel(const FastSpec fs, unsigned long long* out) {
  unsigned long long c = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < fs.dense_cap; i += (uint64_t)gridDim.x * blockDim.x)
    c += fs.dense_tab[i * fs.dense_stride + fs.dense_presence_word] != 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
int launch_dense_count(const FastSpec& fs, unsigned long long* d_out, cudaStream_t s) {
  dense_count_kernel<<<fast_grid(((int64_t)fs.dense_cap + 255) / 256), 256, 0, s>>>(fs, d_out);
  return 1;
}

}  // namespace b200q
