// ASan/UBSan fuzz harness of the host-side decoders (plan_decode.cc, arrow_ipc.cc): reads seed files, mutates, decodes.
// Every outcome must be a value or a PlanError — never a crash, an out-of-bounds read or undefined behaviour.
//   usage: plan_decode_fuzz <seed> <iterations> [--ipc] <seed files...>      (--ipc: seeds are ScalarValue.ipc_bytes literals)
//   build: see tools/fuzz/run.sh
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <vector>
#include "ir.h"
using namespace b200q;
int main(int argc, char** argv) {
  std::vector<std::vector<uint8_t>> seeds;
  bool ipc = false;
  for (int i = 3; i < argc; i++) {
    if (std::string(argv[i]) == "--ipc") { ipc = true; continue; } std::ifstream f(argv[i], std::ios::binary); seeds.emplace_back(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>()); }
  std::mt19937_64 rng(atoll(argv[1])); long n = atol(argv[2]); long ok = 0, err = 0;
  for (long it = 0; it < n; it++) {
    std::vector<uint8_t> b = seeds[rng() % seeds.size()];
    const int k = rng() % 10;
    if (k < 4) { for (int j = 0, m = 1 + rng() % 4; j < m; j++) b[rng() % b.size()] = (uint8_t)rng(); }
    else if (k < 6) b.resize(rng() % b.size());
    if (b.empty()) b.push_back(0);
    else if (k < 8) { size_t i = rng() % b.size(), j = std::min(b.size(), i + 1 + rng() % 16); b.erase(b.begin() + i, b.begin() + j); }
    else { size_t i = rng() % b.size(); for (int j = 0, m = 1 + rng() % 8; j < m; j++) b.insert(b.begin() + i, (uint8_t)rng()); }
    try {
      if (ipc) { ExprP e = decode_ipc_literal(b.data(), b.size()); ok += e != nullptr; }
      else { PlanP p = decode_plan(b.data(), b.size(), (int)(rng() % 2)); std::string s = explain_plan(p); ok += !s.empty(); }
    }
    catch (const std::exception&) { err++; }
  }
  printf("ok=%ld err=%ld\n", ok, err);
}
