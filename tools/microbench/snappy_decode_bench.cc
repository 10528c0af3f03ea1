// Host microbenchmark of the library's Snappy decoder (parquet_meta.cc) on one page body.
// Make the inputs:  python -c "import numpy as np, pyarrow as pa; r=np.random.default_rng(3).integers(0,2_000_000,1<<20,dtype=np.int64).tobytes(); open('/tmp/snappy_case.raw','wb').write(r); open('/tmp/snappy_case.bin','wb').write(pa.Codec('snappy').compress(r, asbytes=True))"
// Build:            g++ -O3 -std=c++17 -Iblaze_b200/csrc -o /tmp/snappy_bench tools/microbench/snappy_decode_bench.cc blaze_b200/csrc/parquet_meta.cc
// Measured in the build container (int64 values < 2 M, 8.4 MB page): byte-wise decoder 13.2 ms (0.64 GB/s) -> block-copy fast loop 8.2 ms (1.02 GB/s).
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>
#include "parquet_meta.h"
using namespace b200q;
int main() {
  std::ifstream f("/tmp/snappy_case.bin", std::ios::binary); std::vector<uint8_t> c((std::istreambuf_iterator<char>(f)), {});
  std::ifstream g("/tmp/snappy_case.raw", std::ios::binary); std::vector<uint8_t> r((std::istreambuf_iterator<char>(g)), {});
  ByteBuf out; double best = 1e9;
  for (int i = 0; i < 20; i++) { out.clear(); auto t0 = std::chrono::steady_clock::now(); snappy_uncompress(c.data(), c.size(), out); double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); if (ms < best) best = ms; }
  printf("%zu -> %zu bytes, %.3f ms = %.2f GB/s, equal=%d\n", c.size(), out.size(), best, out.size() / best / 1e6, (int)(out.size() == r.size() && memcmp(out.data(), r.data(), r.size()) == 0));
}
