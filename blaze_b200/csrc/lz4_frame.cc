// LZ4 frame encoder of the shuffle files' compression blocks (host side; the GPU produces the uncompressed batch_serde bytes).
//
// The reference frames every shuffle block as `u32 LE length ‖ LZ4 frame` with lz4_flex's FrameEncoder
// (datafusion-ext-plans/src/common/ipc_compression.rs:34-112, codec "lz4" = spark.io.compression.codec default :271-283).
// lz4_flex is not vendored; what is restated here is the published LZ4 Frame Format 1.6.x and Block Format: magic
// 0x184D2204, FLG (version 01, independent blocks, no checksums, no content size), BD (4 MiB blocks), HC = second byte of
// xxHash32(descriptor, 0); data blocks `u32 size (bit 31 = stored) ‖ bytes`; EndMark 0.  Any conforming decoder
// (lz4_flex's FrameDecoder, liblz4's LZ4F_decompress via pyarrow in the tests) reads it.  The block compressor is a
// single-pass greedy matcher (64 Ki-entry hash table of 4-byte sequences, 64 KiB window): the byte-plane layout of
// batch_serde makes long runs of equal high-order bytes, which this finds as RLE-style matches.
#include "lz4_frame.h"

#include <cstring>

namespace b200q {

namespace {

inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;

}  // namespace

uint32_t xxhash32(const uint8_t* p, size_t n, uint32_t seed) {
  const uint8_t* end = p + n;
  uint32_t h;
  if (n >= 16) {
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = rotl(v1 + rd32(p) * P2, 13) * P1; v2 = rotl(v2 + rd32(p + 4) * P2, 13) * P1;
      v3 = rotl(v3 + rd32(p + 8) * P2, 13) * P1; v4 = rotl(v4 + rd32(p + 12) * P2, 13) * P1;
      p += 16;
    } while (p + 16 <= end);
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
  } else h = seed + P5;
  h += (uint32_t)n;
  while (p + 4 <= end) { h = rotl(h + rd32(p) * P3, 17) * P4; p += 4; }
  while (p < end) { h = rotl(h + (*p) * P5, 11) * P1; p++; }
  h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
  return h;
}

size_t lz4_block_bound(size_t n) { return n + n / 255 + 16; }

// LZ4 block format: sequences of [token][literal length bytes][literals][offset u16 LE][match length bytes]; the last
// sequence has literals only; the last 5 bytes are literals and no match starts within the last 12 bytes.
size_t lz4_block_compress(const uint8_t* src, size_t n, uint8_t* dst) {
  constexpr int HBITS = 16;
  static thread_local uint32_t table[1 << HBITS];
  uint8_t* op = dst;
  size_t anchor = 0;
  auto emit_literals_and_match = [&](size_t lit_len, const uint8_t* lit, bool has_match, size_t offset, size_t mlen) {
    uint8_t* token = op++;
    size_t l = lit_len;
    if (l >= 15) { *token = 0xF0; l -= 15; while (l >= 255) { *op++ = 255; l -= 255; } *op++ = (uint8_t)l; } else *token = (uint8_t)(l << 4);
    memcpy(op, lit, lit_len); op += lit_len;
    if (!has_match) return;
    *op++ = (uint8_t)offset; *op++ = (uint8_t)(offset >> 8);
    size_t m = mlen - 4;
    if (m >= 15) { *token |= 15; m -= 15; while (m >= 255) { *op++ = 255; m -= 255; } *op++ = (uint8_t)m; } else *token |= (uint8_t)m;
  };
  if (n >= 13) {
    memset(table, 0, sizeof(table));                                  // positions are stored + 1 (0 = empty)
    const size_t mflimit = n - 12, matchlimit = n - 5;
    size_t ip = 0, misses = 0;
    while (ip < mflimit) {
      const uint32_t seq = rd32(src + ip);
      const uint32_t h = (seq * 2654435761u) >> (32 - HBITS);
      const size_t ref = table[h];
      table[h] = (uint32_t)(ip + 1);
      if (ref && ip + 1 - ref <= 65535 && rd32(src + ref - 1) == seq) {
        const size_t r = ref - 1;
        size_t m = 4;
        while (ip + m + 8 <= matchlimit && rd64(src + ip + m) == rd64(src + r + m)) m += 8;
        while (ip + m < matchlimit && src[ip + m] == src[r + m]) m++;
        emit_literals_and_match(ip - anchor, src + anchor, true, ip - r, m);
        ip += m; anchor = ip; misses = 0;
      } else {
        ip += 1 + (misses++ >> 6);                                    // skip faster through incompressible data
      }
    }
  }
  emit_literals_and_match(n - anchor, src + anchor, false, 0, 0);
  return (size_t)(op - dst);
}

void lz4_frame_append(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  constexpr size_t BLOCK = 4u << 20;
  const uint8_t desc[2] = {0x60, 0x70};                               // FLG: version 01, block independence; BD: 4 MiB
  const uint8_t hdr[7] = {0x04, 0x22, 0x4D, 0x18, desc[0], desc[1], (uint8_t)(xxhash32(desc, 2, 0) >> 8)};
  out.insert(out.end(), hdr, hdr + 7);
  for (size_t pos = 0; pos < n; pos += BLOCK) {
    const size_t len = n - pos < BLOCK ? n - pos : BLOCK;
    const size_t at = out.size();
    out.resize(at + 4 + lz4_block_bound(len));
    size_t clen = lz4_block_compress(src + pos, len, out.data() + at + 4);
    uint32_t word;
    if (clen >= len) { memcpy(out.data() + at + 4, src + pos, len); clen = len; word = (uint32_t)len | 0x80000000u; }     // stored block
    else word = (uint32_t)clen;
    memcpy(out.data() + at, &word, 4);
    out.resize(at + 4 + clen);
  }
  const uint8_t endmark[4] = {0, 0, 0, 0};
  out.insert(out.end(), endmark, endmark + 4);
}

}  // namespace b200q
