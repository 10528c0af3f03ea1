// Device-side pieces shared by the generic (kernels.cu) and the specialised (kernels_fast.cu) HashAgg
// update kernels: atomics, the key hash, and the open-addressing find-or-insert protocol.
//
// Slot protocol (cf. the reference's 8-lane groups of {u32 hash, u32 record id}, agg_hash_map.rs:77-136):
//   tag 0 = empty, 1 = locked (an inserting thread is writing the key), else 0x80000000|fingerprint.
//   insert: CAS tag 0->1, write key words + accumulator identities + flags, fence, release-store the tag.
//   lookup: one relaxed (L2-coherent) load of the header; key words are compared only on a tag match.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace b200q {

constexpr unsigned TAG_EMPTY = 0, TAG_LOCKED = 1;
constexpr unsigned FLAG_SLOT_LOCK = 1u << 15;     // guards 128-bit min/max updates

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ ulonglong2 ld_relaxed_v2u64(const unsigned long long* p) {
  ulonglong2 v; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_add_f64(unsigned long long* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void red_min_s64(unsigned long long* p, long long v) { asm volatile("red.global.min.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_max_s64(unsigned long long* p, long long v) { asm volatile("red.global.max.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
constexpr uint64_t AGG_HASH_SEED = 0x9E3779B97F4A7C15ULL;
__device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// hash of the key words (NULL keys canonicalised to 0) and the key-is-NULL mask; not observable in results.
// One 64-bit finaliser per key: extra words / the null mask are folded in with odd multipliers first.
__device__ __forceinline__ uint64_t agg_hash2(uint64_t k0, uint64_t k1, uint32_t knull) {
  return mix64((AGG_HASH_SEED ^ k0) + k1 * 0xD6E8FEB86659FD93ULL + (uint64_t)knull * 0xA0761D6478BD642FULL);
}
__device__ __forceinline__ uint64_t agg_hash_words(const uint64_t* kw, int nkw, uint32_t knull) {
  if (nkw <= 2) return agg_hash2(nkw > 0 ? kw[0] : 0, nkw > 1 ? kw[1] : 0, knull);
  uint64_t x = AGG_HASH_SEED ^ kw[0];
  for (int i = 1; i < nkw; i++) x = x * 0xD6E8FEB86659FD93ULL + kw[i];
  return mix64(x + (uint64_t)knull * 0xA0761D6478BD642FULL);
}

// set an accumulator-valid bit in the key entry's header (only when the copy we saw did not have it yet)
__device__ __forceinline__ void slot_mark(unsigned long long* key_entry, unsigned flags_seen, int vbit) {
  if (vbit != 0xFF && !((flags_seen >> vbit) & 1)) atomicOr((unsigned*)key_entry + 1, 1u << vbit);
}

// capacity < 2^32 slots: the slot comes from the high 32 hash bits with one 32-bit multiply-high, the tag from the low bits
__device__ __forceinline__ uint64_t agg_first_slot(uint64_t h, uint64_t capacity) { return __umulhi((unsigned)(h >> 32), (unsigned)capacity); }
__device__ __forceinline__ unsigned agg_tag(uint64_t h) { return (unsigned)h | 0x80000000u; }
__device__ __forceinline__ uint64_t agg_next_slot(uint64_t s, uint64_t capacity) { return s + 1 == capacity ? 0 : s + 1; }
constexpr uint64_t AGG_NO_SLOT = ~0ULL;

// Returns the slot number of the key (inserting it if absent) or AGG_NO_SLOT when the table is at its load limit
// and the row must be deferred.  *flags = the slot's flags word as last seen (may be stale: only used to skip
// redundant validity marking).
__device__ __forceinline__ uint64_t agg_find_or_insert(const AggLayout& lay, const AggTable& tab, const uint64_t* kw, uint32_t knull,
                                                       uint64_t h, unsigned* flags_out, bool* inserted) {
  const unsigned tag = agg_tag(h);
  uint64_t s = agg_first_slot(h, tab.capacity);
  while (true) {
    unsigned long long* ke = tab.keys + s * (uint64_t)lay.kstride;
    const unsigned long long hdr = ld_relaxed_u64(ke);
    const unsigned t = (unsigned)hdr;
    unsigned flags = (unsigned)(hdr >> 32);
    if (t == tag) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");                          // pairs with the inserter's st.release: key words / identities are visible before they are read or RED-ed
      bool eq = (flags >> 16) == knull;
      for (int i = 0; eq && i < lay.nkw; i++) eq = ld_relaxed_u64(ke + 1 + i) == kw[i];
      if (eq) { *flags_out = flags; return s; }
    } else if (t == TAG_EMPTY) {
      if (ld_relaxed_u64(tab.counters) >= tab.max_groups) return AGG_NO_SLOT;   // at the load limit: defer the row
      if (atomicCAS((unsigned*)ke, TAG_EMPTY, TAG_LOCKED) == TAG_EMPTY) {
        for (int i = 0; i < lay.nkw; i++) ke[1 + i] = kw[i];
        unsigned long long* ae = tab.accs + s * (uint64_t)lay.astride;
        for (int i = 0; i < lay.astride; i++) ae[i] = lay.init[i];
        flags = lay.init_flags | (knull << 16);
        ((unsigned*)ke)[1] = flags;
        st_release_u32((unsigned*)ke, tag);                                       // release: key words, flags and accumulator identities first
        *inserted = true;            // the caller adds to tab.counters[0] (warp-aggregated where it can)
        *flags_out = flags;
        return s;
      }
      continue;                                                                 // lost the race: look at the same slot again
    } else if (t == TAG_LOCKED) {
      continue;                                                                 // being published by another thread
    }
    s = agg_next_slot(s, tab.capacity);
  }
}

}  // namespace b200q
