// LZ4 frame encoder (lz4_frame.cc): the compression blocks of the shuffle files.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace b200q {

uint32_t xxhash32(const uint8_t* p, size_t n, uint32_t seed);
size_t lz4_block_bound(size_t n);
size_t lz4_block_compress(const uint8_t* src, size_t n, uint8_t* dst);          // dst holds lz4_block_bound(n) bytes
void lz4_frame_append(const uint8_t* src, size_t n, std::vector<uint8_t>& out);  // one complete frame appended to `out`

}  // namespace b200q
