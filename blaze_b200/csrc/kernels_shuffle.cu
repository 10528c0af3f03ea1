// ShuffleWriterExec on the GPU: hash partition + `batch_serde` encode of one resident batch.
//
// Replaces, for HashPartitioning / single-partition outputs over fixed-width columns:
//   evaluate_hashes + evaluate_partition_ids     datafusion-ext-plans/src/shuffle/mod.rs:163-188
//   sort_batches_by_partition_id                 datafusion-ext-plans/src/shuffle/buffered_data.rs:284-351
//   PartitionedBatchesIterator + write_batch     buffered_data.rs:219-282, datafusion-ext-commons/src/io/batch_serde.rs:66-77,264-306
// The reference sorts (part_id, batch, row) triples on the host, interleaves the rows into a partition-sorted batch
// (a full copy) and then transposes every column of every sub-batch into byte planes (a second copy).  Here one pass
// computes the partition ids and their histogram, a one-CTA pass lays the output out (every record's size follows from
// the counts alone), and ONE pass moves the data: a CTA ranks a 4096-row tile per partition in shared memory, reserves
// the tile's rows of every partition with one global atomic per (tile, partition), stages each column through shared
// memory in partition order and writes the byte planes of the final wire format directly — rows that are neighbours
// in a partition are neighbours in every plane, so a warp's byte stores fall on one or two 32-byte sectors.
// Traffic: keys once more for the ids (+2 B/row of ids), every column read once, every encoded byte written once.
// The row order inside a partition is not a contract (the reference's radix sort is unstable, rdx_sort.rs:55-73).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

#include "kernels_shuffle.cuh"
#include "tma.cuh"

namespace b200q {

namespace {

__device__ __forceinline__ uint32_t sh_rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ __forceinline__ uint32_t sh_mix_k1(uint32_t k1) { k1 *= 0xcc9e2d51u; k1 = sh_rotl32(k1, 15); k1 *= 0x1b873593u; return k1; }
__device__ __forceinline__ uint32_t sh_mix_h1(uint32_t h1, uint32_t k1) { h1 ^= k1; h1 = sh_rotl32(h1, 13); return h1 * 5 + 0xe6546b64u; }
__device__ __forceinline__ uint32_t sh_fmix(uint32_t h1, uint32_t len) { h1 ^= len; h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16; return h1; }

int sm_count() {
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

constexpr int PID_BLOCK = 256;

// Spark murmur3 (seed 42) chained over the key columns, NULL leaves the running hash unchanged (hash/mur.rs:19-87,
// spark_hash.rs:84-200), pmod P (shuffle/mod.rs:178-188); histogram in shared memory, one global atomic per (CTA, partition)
__global__ void __launch_bounds__(PID_BLOCK) shuffle_pids_kernel(const ShufSpec sp, long long n, uint16_t* __restrict__ pids, unsigned long long* counts) {
  __shared__ unsigned s_hist[SHUF_MAX_PARTS];
  const int P = sp.num_partitions;
  for (int p = threadIdx.x; p < P; p += PID_BLOCK) s_hist[p] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)PID_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * PID_BLOCK) {
    uint32_t h = 42;
    for (int c = 0; c < sp.nkeys; c++) {
      const ShufCol& col = sp.col[sp.key_col[c]];
      if (col.validity) { const unsigned long long bi = (unsigned long long)i + col.bit_offset; if (!((col.validity[bi >> 3] >> (bi & 7)) & 1)) continue; }
      uint32_t w[4]; int nw;
      switch (sp.key_phys[c]) {
        case PH_BOOL: { const unsigned long long bi = (unsigned long long)i + col.bit_offset; w[0] = (((const uint8_t*)col.values)[bi >> 3] >> (bi & 7)) & 1; nw = 1; break; }
        case PH_I8: w[0] = (uint32_t)(int32_t)((const int8_t*)col.values)[i]; nw = 1; break;
        case PH_I16: w[0] = (uint32_t)(int32_t)((const int16_t*)col.values)[i]; nw = 1; break;
        case PH_I32: case PH_F32: w[0] = ((const uint32_t*)col.values)[i]; nw = 1; break;
        case PH_I64: case PH_F64: { const unsigned long long v = ((const unsigned long long*)col.values)[i]; w[0] = (uint32_t)v; w[1] = (uint32_t)(v >> 32); nw = 2; break; }
        default: { const unsigned long long a = ((const unsigned long long*)col.values)[2 * i], b = ((const unsigned long long*)col.values)[2 * i + 1];
                   w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32); w[2] = (uint32_t)b; w[3] = (uint32_t)(b >> 32); nw = 4; break; }
      }
      uint32_t h1 = h;
      for (int k = 0; k < nw; k++) h1 = sh_mix_h1(h1, sh_mix_k1(w[k]));
      h = sh_fmix(h1, (uint32_t)(4 * nw));
    }
    int32_t m = (int32_t)h % P;                                                      // rem_euclid
    if (m < 0) m += P;
    pids[i] = (uint16_t)m;
    atomicAdd(&s_hist[m], 1u);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += PID_BLOCK) if (s_hist[p]) atomicAdd(counts + p, (unsigned long long)s_hist[p]);
}

// part_off = exclusive prefix of the partitions' encoded sizes; cursors = 0
__global__ void __launch_bounds__(1024) shuffle_layout_kernel(const ShufSpec sp, const unsigned long long* __restrict__ counts, unsigned long long* part_off, unsigned long long* cursors) {
  __shared__ unsigned long long s[SHUF_MAX_PARTS + 1];
  const int P = sp.num_partitions;
  for (int p = threadIdx.x; p < P; p += 1024) { s[p] = shuf_partition_bytes(sp, counts[p]); cursors[p] = 0; }
  __syncthreads();
  if (threadIdx.x == 0) { unsigned long long acc = 0; for (int p = 0; p < P; p++) { const unsigned long long b = s[p]; s[p] = acc; acc += b; } s[P] = acc; }
  __syncthreads();
  for (int p = threadIdx.x; p <= P; p += 1024) part_off[p] = s[p];
}

__device__ __forceinline__ unsigned long long col_offset(const ShufSpec& sp, int c, unsigned long long m, unsigned long long m8, uint32_t vl) {
  return (unsigned long long)vl + (unsigned long long)c + (unsigned long long)sp.col[c].k8 * m8 + (unsigned long long)sp.col[c].kw * m;
}

// record headers: varint(m) and the per-column `has null buffer` byte (io/mod.rs:60-69, batch_serde.rs:274-284); one CTA per partition
__global__ void __launch_bounds__(128) shuffle_headers_kernel(const ShufSpec sp, const unsigned long long* __restrict__ counts, const unsigned long long* __restrict__ part_off, uint8_t* out) {
  const int p = blockIdx.x;
  const unsigned long long t = counts[p], B = (unsigned long long)sp.batch_size;
  if (t == 0) return;
  const unsigned long long nrec = (t + B - 1) / B, F = shuf_record_bytes(sp, B);
  for (unsigned long long rec = threadIdx.x; rec < nrec; rec += 128) {
    const unsigned long long m = rec == nrec - 1 ? t - rec * B : B, m8 = (m + 7) >> 3;
    uint8_t* base = out + part_off[p] + rec * F;
    unsigned long long v = m; uint32_t vl = 0;
    while (v >= 128) { base[vl++] = (uint8_t)(128 + (v & 127)); v >>= 7; }
    base[vl++] = (uint8_t)v;
    for (int c = 0; c < sp.ncols; c++) base[col_offset(sp, c, m, m8, vl)] = sp.col[c].nullable ? 1 : 0;
  }
}

__device__ __forceinline__ void or_bit(uint8_t* out, unsigned long long byte_off, unsigned bit) {
  const unsigned long long a = (unsigned long long)(uintptr_t)out + byte_off;
  atomicOr((unsigned*)(uintptr_t)(a & ~3ull), 1u << (((unsigned)(a & 3ull) << 3) + bit));
}

constexpr int ENC_NT = 512, ENC_RPT = SHUF_TILE / ENC_NT, SHUF_SMEM_PARTS = 512;

// ---- TMA (cp.async.bulk) staging of the input columns -------------------------------------------------------------------
// One elected thread copies a whole tile of a column (4096 values, contiguous in HBM) into shared memory with ONE bulk copy
// that completes on an mbarrier; two buffers, so the copy of the next column (or of the next tile's first column) is in
// flight while the current one is permuted and written out, and the first copy of a tile overlaps its ranking phase.

// where sorted position `i` of the tile lands: byte offset of its record + its row inside the record; rows of that record
__device__ __forceinline__ unsigned long long dest_of(const ShufSpec& sp, unsigned long long F, unsigned B, unsigned p, unsigned idx, const unsigned long long* counts,
                                                       const unsigned long long* part_off, unsigned& m, unsigned& j) {
  const unsigned t = (unsigned)counts[p], rec = idx / B, nrec = (t + B - 1) / B;
  j = idx - rec * B;
  m = rec == nrec - 1 ? t - rec * B : B;
  return part_off[p] + (unsigned long long)rec * F;
}

template <bool TMA>
__global__ void __launch_bounds__(ENC_NT, 2) shuffle_encode_kernel(const ShufSpec sp, const uint16_t* __restrict__ pids, long long n, const unsigned long long* __restrict__ counts_g,
                                                                   const unsigned long long* __restrict__ part_off_g, unsigned long long* cursors, uint8_t* out) {
#ifdef B200Q_EMULATED_DEVICE                                                 // tools/emu: blocks run one at a time, shared memory is a static array
  static unsigned long long smem_words[(SHUF_TILE * 26 + SHUF_MAX_PARTS * 16) / 8];
  unsigned char* smem = (unsigned char*)smem_words;
#else
  extern __shared__ __align__(16) unsigned char smem[];
#endif
  const int P = sp.num_partitions;
  unsigned long long* s_in = (unsigned long long*)smem;                     // TMA: two input buffers of SHUF_TILE values (input order)
  unsigned long long* s_val = s_in + (TMA ? 2 * SHUF_TILE : 0);             // SHUF_TILE values in partition order
  unsigned long long* s_gbase = s_val + SHUF_TILE;                          // P: index inside the partition of the tile's first row of it
  unsigned* s_cnt = (unsigned*)(s_gbase + P);                               // P: rows of the tile per partition
  unsigned* s_start = s_cnt + P;                                            // P: exclusive prefix of s_cnt
  uint16_t* s_p = (uint16_t*)(s_start + P);                                 // SHUF_TILE: partition of every sorted position
  __shared__ unsigned s_warp[ENC_NT / 32];
  const unsigned B = (unsigned)sp.batch_size, B8 = (B + 7) >> 3;
  const unsigned long long F = shuf_record_bytes(sp, B);
  const uint32_t vlB = shuf_varint_len(B);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ntiles = (n + SHUF_TILE - 1) / SHUF_TILE;
  // rows / byte offset of every partition are looked up once per row and tile: keep them in shared memory when they fit
  __shared__ unsigned long long s_tab[2 * SHUF_SMEM_PARTS];
  if (P <= SHUF_SMEM_PARTS) {
    for (int p = tid; p < P; p += ENC_NT) { s_tab[p] = counts_g[p]; s_tab[SHUF_SMEM_PARTS + p] = part_off_g[p]; }
    __syncthreads();
  }
  const unsigned long long* counts = P <= SHUF_SMEM_PARTS ? s_tab : counts_g;
  const unsigned long long* part_off = P <= SHUF_SMEM_PARTS ? s_tab + SHUF_SMEM_PARTS : part_off_g;
  // TMA load sequence of this CTA: load j = column s_ec[j % ne] of the CTA's (j / ne)-th tile, into buffer j & 1
  __shared__ unsigned long long s_bar[2];                               // 8-byte aligned by type
  __shared__ unsigned char s_ec[SHUF_MAX_COLS];
  __shared__ int s_ne;
  unsigned consumed = 0;
  auto issue = [&](unsigned j) {
    if (!TMA || s_ne == 0) return;
    const long long tl = blockIdx.x + (long long)(j / (unsigned)s_ne) * gridDim.x;
    if (tl >= ntiles || (tl + 1) * SHUF_TILE > n) return;                   // only full tiles are staged by bulk copies
    const int c = s_ec[j % (unsigned)s_ne];
    const unsigned w = sp.col[c].width, bytes = SHUF_TILE * w;
    mbar_expect_tx(&s_bar[j & 1], bytes);
    bulk_g2s(s_in + (size_t)(j & 1) * SHUF_TILE, (const uint8_t*)sp.col[c].values + (size_t)tl * SHUF_TILE * w, bytes, &s_bar[j & 1]);
  };
  if (TMA) {
    if (tid == 0) {
      int ne = 0;
      for (int c = 0; c < sp.ncols; c++) if (sp.col[c].tma) s_ec[ne++] = (unsigned char)c;
      s_ne = ne;
      mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) { issue(0); issue(1); }
  }
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long t0 = tile * SHUF_TILE;
    const int rows = (int)min((long long)SHUF_TILE, n - t0);
    for (int p = tid; p < P; p += ENC_NT) s_cnt[p] = 0;
    __syncthreads();
    unsigned lpos[ENC_RPT];                                                 // rank inside (tile, partition), then position in the tile's partition order
    {
      unsigned pid[ENC_RPT];
#pragma unroll
      for (int k = 0; k < ENC_RPT; k++) {
        const int i = k * ENC_NT + tid;
        pid[k] = 0; lpos[k] = 0;
        if (i < rows) { pid[k] = pids ? pids[t0 + i] : 0; lpos[k] = atomicAdd(&s_cnt[pid[k]], 1u); }
      }
      __syncthreads();
      {   // exclusive scan of s_cnt, one contiguous span of partitions per thread; reserve the tile's rows of every partition
        const int per = (P + ENC_NT - 1) / ENC_NT, lo = min(P, tid * per), hi = min(P, lo + per);
        unsigned sum = 0;
        for (int p = lo; p < hi; p++) sum += s_cnt[p];
        unsigned inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned wbase = 0;
        for (int w = 0; w < warp; w++) wbase += s_warp[w];
        unsigned run = wbase + inc - sum;
        for (int p = lo; p < hi; p++) {
          const unsigned c = s_cnt[p];
          s_start[p] = run; run += c;
          if (c) s_gbase[p] = atomicAdd(cursors + p, (unsigned long long)c);
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < ENC_RPT; k++) {
        const int i = k * ENC_NT + tid;
        if (i < rows) { lpos[k] += s_start[pid[k]]; s_p[lpos[k]] = (uint16_t)pid[k]; }
      }
    }
    __syncthreads();
    // destination of the sorted positions this thread writes out: record base + row inside the record; bit k of `full`: the
    // record holds batch_size rows (every record but a partition's last one) -> its plane stride and column offsets are constants
    unsigned long long wb[ENC_RPT]; unsigned full = 0;
#pragma unroll
    for (int k = 0; k < ENC_RPT; k++) {
      const int i = k * ENC_NT + tid;
      wb[k] = 0;
      if (i < rows) {
        const unsigned p = s_p[i];
        unsigned m, j;
        wb[k] = dest_of(sp, F, B, p, (unsigned)s_gbase[p] + ((unsigned)i - s_start[p]), counts, part_off, m, j) + j;
        if (m == B) full |= 1u << k;
      }
    }
    for (int c = 0; c < sp.ncols; c++) {
      const unsigned width = sp.col[c].width, nullable = sp.col[c].nullable;
      const void* __restrict__ values = sp.col[c].values;
      const int nh = width == 16 ? 2 : (width ? 1 : 0);
      const int nb = width < 8 ? (int)width : 8;
      // data region of column c inside a full record
      const unsigned long long off_full = (unsigned long long)vlB + (unsigned)c + (unsigned long long)sp.col[c].k8 * B8 + (unsigned long long)sp.col[c].kw * B + 1 + (nullable ? B8 : 0);
      for (int h = 0; h < nh; h++) {
        const bool staged = TMA && sp.col[c].tma && rows == SHUF_TILE;
        if (staged) {                                                       // the column's tile is (being) copied into s_in[consumed & 1]
          mbar_wait(&s_bar[consumed & 1], (consumed >> 1) & 1);
          const unsigned long long* in = s_in + (size_t)(consumed & 1) * SHUF_TILE;
          if (width == 8) {
#pragma unroll
            for (int k = 0; k < ENC_RPT; k++) s_val[lpos[k]] = in[k * ENC_NT + tid];
          } else if (width == 4) {
#pragma unroll
            for (int k = 0; k < ENC_RPT; k++) s_val[lpos[k]] = ((const uint32_t*)in)[k * ENC_NT + tid];
          } else if (width == 2) {
#pragma unroll
            for (int k = 0; k < ENC_RPT; k++) s_val[lpos[k]] = ((const uint16_t*)in)[k * ENC_NT + tid];
          } else {
#pragma unroll
            for (int k = 0; k < ENC_RPT; k++) s_val[lpos[k]] = ((const uint8_t*)in)[k * ENC_NT + tid];
          }
        } else if (width == 8) {
#pragma unroll
          for (int k = 0; k < ENC_RPT; k++) { const int i = k * ENC_NT + tid; if (i < rows) s_val[lpos[k]] = ((const unsigned long long*)values)[t0 + i]; }
        } else if (width == 4) {
#pragma unroll
          for (int k = 0; k < ENC_RPT; k++) { const int i = k * ENC_NT + tid; if (i < rows) s_val[lpos[k]] = ((const uint32_t*)values)[t0 + i]; }
        } else if (width == 16) {
#pragma unroll
          for (int k = 0; k < ENC_RPT; k++) { const int i = k * ENC_NT + tid; if (i < rows) s_val[lpos[k]] = ((const unsigned long long*)values)[2 * (t0 + i) + h]; }
        } else if (width == 2) {
#pragma unroll
          for (int k = 0; k < ENC_RPT; k++) { const int i = k * ENC_NT + tid; if (i < rows) s_val[lpos[k]] = ((const uint16_t*)values)[t0 + i]; }
        } else {
#pragma unroll
          for (int k = 0; k < ENC_RPT; k++) { const int i = k * ENC_NT + tid; if (i < rows) s_val[lpos[k]] = ((const uint8_t*)values)[t0 + i]; }
        }
        __syncthreads();
        if (staged) { if (tid == 0) issue(consumed + 2); consumed++; }     // the buffer just read is free: start the copy two loads ahead
#pragma unroll
        for (int k = 0; k < ENC_RPT; k++) {
          const int i = k * ENC_NT + tid;
          if (i < rows) {
            const unsigned long long v = s_val[i];
            if (full & (1u << k)) {
              uint8_t* a = out + wb[k] + off_full + (unsigned long long)(h * 8) * B;
              if (nb == 8) {
#pragma unroll
                for (int b = 0; b < 8; b++) { *a = (uint8_t)(v >> (8 * b)); a += B; }
              } else if (nb == 4) {
#pragma unroll
                for (int b = 0; b < 4; b++) { *a = (uint8_t)(v >> (8 * b)); a += B; }
              } else {
                for (int b = 0; b < nb; b++) { *a = (uint8_t)(v >> (8 * b)); a += B; }
              }
            } else {                                                        // the short last record of a partition
              const unsigned p = s_p[i];
              unsigned m, j;
              dest_of(sp, F, B, p, (unsigned)s_gbase[p] + ((unsigned)i - s_start[p]), counts, part_off, m, j);
              const unsigned long long m8 = (m + 7) >> 3;
              uint8_t* a = out + wb[k] + col_offset(sp, c, m, m8, shuf_varint_len(m)) + 1 + (nullable ? m8 : 0) + (unsigned long long)(h * 8) * m;
              for (int b = 0; b < nb; b++) a[(size_t)b * m] = (uint8_t)(v >> (8 * b));
            }
          }
        }
        __syncthreads();
      }
      if (nullable || width == 0) {
        // bit regions (validity bitmaps, Boolean values): set bits are OR-ed into the zeroed buffer from the rows' own destinations
        const uint8_t* __restrict__ validity = sp.col[c].validity;
        const unsigned bit_offset = sp.col[c].bit_offset;
#pragma unroll
        for (int k = 0; k < ENC_RPT; k++) {
          const int i = k * ENC_NT + tid;
          if (i < rows) {
            const long long r = t0 + i;
            const unsigned p = s_p[lpos[k]];
            unsigned m, j;
            const unsigned long long rb = dest_of(sp, F, B, p, (unsigned)s_gbase[p] + (lpos[k] - s_start[p]), counts, part_off, m, j);
            const unsigned long long m8 = (m + 7) >> 3;
            const unsigned long long cb = rb + col_offset(sp, c, m, m8, shuf_varint_len(m)) + 1;
            const unsigned long long bi = (unsigned long long)r + bit_offset;
            if (nullable) {
              const bool valid = validity ? ((validity[bi >> 3] >> (bi & 7)) & 1) : true;
              if (valid) or_bit(out, cb + (j >> 3), j & 7);
            }
            if (width == 0 && ((((const uint8_t*)values)[bi >> 3] >> (bi & 7)) & 1)) or_bit(out, cb + (nullable ? m8 : 0) + (j >> 3), j & 7);
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

int launch_shuffle_pids(const ShufSpec& sp, int64_t n, uint16_t* d_pids, unsigned long long* d_counts, cudaStream_t s) {
  if (n <= 0) return 0;
  const int64_t want = (n + PID_BLOCK * 8 - 1) / (PID_BLOCK * 8);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8));
  shuffle_pids_kernel<<<grid, PID_BLOCK, 0, s>>>(sp, n, d_pids, d_counts);
  return 1;
}

int launch_shuffle_layout(const ShufSpec& sp, const unsigned long long* d_counts, unsigned long long* d_part_off, unsigned long long* d_cursors, uint8_t* d_out, cudaStream_t s) {
  shuffle_layout_kernel<<<1, 1024, 0, s>>>(sp, d_counts, d_part_off, d_cursors);
  if (!d_out) return 1;
  shuffle_headers_kernel<<<sp.num_partitions, 128, 0, s>>>(sp, d_counts, d_part_off, d_out);
  return 2;
}

int launch_shuffle_encode(const ShufSpec& sp, const uint16_t* d_pids, int64_t n, const unsigned long long* d_counts, const unsigned long long* d_part_off,
                          unsigned long long* d_cursors, uint8_t* d_out, cudaStream_t s) {
  if (n <= 0) return 0;
  ShufSpec spx = sp;
  bool tma = false;
#ifndef B200Q_EMULATED_DEVICE
  // measured on B200 (profiles/r02_shapes_shuffle_*): the bulk-copy staging costs 64 KB more shared memory per CTA (a smaller L1 for the
  // write-combining of the byte stores) and the kernel is bound by its store phase, not by the column loads: off unless asked for
  static const bool no_tma = getenv("B200Q_SHUFFLE_TMA") == nullptr;
  for (int c = 0; c < spx.ncols; c++) {                                                     // bulk copies need 16-byte aligned sources
    ShufCol& col = spx.col[c];
    col.tma = !no_tma && (col.width == 1 || col.width == 2 || col.width == 4 || col.width == 8) && ((uintptr_t)col.values & 15) == 0 && n >= SHUF_TILE;
    tma = tma || col.tma;
  }
#endif
  const size_t smem = (size_t)SHUF_TILE * (tma ? 24 : 8) + (size_t)spx.num_partitions * 16 + (size_t)SHUF_TILE * 2;
  const int64_t ntiles = (n + SHUF_TILE - 1) / SHUF_TILE;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)sm_count() * 2));
#ifndef B200Q_EMULATED_DEVICE
  cudaFuncSetAttribute(shuffle_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHUF_TILE * 10 + SHUF_MAX_PARTS * 16);      // per device: cheap, idempotent
  cudaFuncSetAttribute(shuffle_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SHUF_TILE * 26 + SHUF_MAX_PARTS * 16);
  if (tma) { shuffle_encode_kernel<true><<<grid, ENC_NT, smem, s>>>(spx, d_pids, n, d_counts, d_part_off, d_cursors, d_out); return 1; }
#endif
  shuffle_encode_kernel<false><<<grid, ENC_NT, smem, s>>>(spx, d_pids, n, d_counts, d_part_off, d_cursors, d_out);
  return 1;
}

}  // namespace b200q
