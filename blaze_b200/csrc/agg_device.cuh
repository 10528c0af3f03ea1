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

__device__ __forceinline__ void slot_mark(unsigned long long* slot, unsigned flags_seen, int vbit) {
  if (vbit != 0xFF && !((flags_seen >> vbit) & 1)) atomicOr((unsigned*)slot + 1, 1u << vbit);
}

// Returns the slot of the key (inserting it if absent) or nullptr when the table is at its load limit and
// the row must be deferred.  *flags = the slot's flags word as last seen (may be stale: only used to skip
// redundant validity marking).  `s` = first slot index to probe.
__device__ __forceinline__ unsigned long long* agg_find_or_insert(const AggLayout& lay, const AggTable& tab, const uint64_t* kw, uint32_t knull,
                                                                   uint64_t h, unsigned* flags_out, bool* inserted) {
  const unsigned tag = (unsigned)(h >> 32) | 0x80000000u;
  uint64_t s = h & tab.mask;
  while (true) {
    unsigned long long* slot = tab.slots + s * (uint64_t)lay.slot_words;
    const unsigned long long hdr = ld_relaxed_u64(slot);
    const unsigned t = (unsigned)hdr;
    unsigned flags = (unsigned)(hdr >> 32);
    if (t == tag) {
      bool eq = (flags >> 16) == knull;
      for (int i = 0; eq && i < lay.nkw; i++) eq = ld_relaxed_u64(slot + 1 + i) == kw[i];
      if (eq) { *flags_out = flags; return slot; }
    } else if (t == TAG_EMPTY) {
      if (ld_relaxed_u64(tab.counters) >= tab.max_groups) return nullptr;       // at the load limit: defer the row
      if (atomicCAS((unsigned*)slot, TAG_EMPTY, TAG_LOCKED) == TAG_EMPTY) {
        for (int i = 0; i < lay.nkw; i++) slot[1 + i] = kw[i];
        for (int i = 1 + lay.nkw; i < lay.slot_words; i++) slot[i] = lay.init[i];
        flags = lay.init_flags | (knull << 16);
        ((unsigned*)slot)[1] = flags;
        __threadfence();
        st_release_u32((unsigned*)slot, tag);
        *inserted = true;            // the caller adds to tab.counters[0] (warp-aggregated where it can)
        *flags_out = flags;
        return slot;
      }
      continue;                                                                 // lost the race: look at the same slot again
    } else if (t == TAG_LOCKED) {
      continue;                                                                 // being published by another thread
    }
    s = (s + 1) & tab.mask;
  }
}

}  // namespace b200q
