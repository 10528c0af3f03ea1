// Thread-per-lane emulation tests of the specialised HashAgg update kernels (tools/emu/README in build_emu.py).
// Every kernel form is run on seeded data (tiny / mid / sparse key ranges, typed + NULLs or lean, filters, keys that
// leave the dense range) and the resulting dense + hash tables are compared, group by group, with a plain host loop.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#include "kernels_tile.cu"   // the tile form the dispatcher of kernels_fast.cu reaches (same translation)
#include "kernels_fast.cu"   // translated copy produced by build_emu.py (inline PTX -> host helpers, <<<>>> -> emu::Launcher)

using namespace b200q;

namespace {

struct Data {
  long long n;
  std::vector<long long> k0, k1, v, w, f;                       // 64-bit host copies of the logical values
  std::vector<uint8_t> nk0, nk1, nv, nw, nf;                    // 1 = NULL
  // physical buffers handed to the kernels
  std::vector<long long> b_k0, b_k1, b_v, b_w, b_f;             // lean (int64)
  std::vector<int32_t> t_k0, t_w; std::vector<int16_t> t_k1; std::vector<int8_t> t_f;   // typed
  std::vector<uint8_t> vb[5];                                   // validity bitmaps (typed)
};

std::vector<uint8_t> bitmap(const std::vector<uint8_t>& nulls) {
  std::vector<uint8_t> b((nulls.size() + 7) / 8 + 8, 0);
  for (size_t i = 0; i < nulls.size(); i++) if (!nulls[i]) b[i >> 3] |= (uint8_t)(1u << (i & 7));
  return b;
}

Data make_data(long long n, long long r0, long long r1, bool typed, unsigned seed) {
  std::mt19937_64 rng(seed); Data d; d.n = n;
  auto rnd = [&](long long lo, long long hi) { return lo + (long long)(rng() % (unsigned long long)(hi - lo)); };
  for (long long i = 0; i < n; i++) {
    long long a = rnd(-5, -5 + r0), b = rnd(100, 100 + r1);
    if (i > n * 2 / 3 && rng() % 5 == 0) a += r0 * (1 + (long long)(rng() % 3));        // leaves the dense range decided on the "first batch"
    if (i > n * 2 / 3 && rng() % 11 == 0) b -= 2 * r1 + 3;
    if (typed) { a = (int32_t)a; b = (int16_t)b; }              // typed keys are stored as int32 / int16: the logical value is the stored one
    d.k0.push_back(a); d.k1.push_back(b);
    d.v.push_back((long long)rng() >> 8); d.w.push_back(rnd(-1000, 1000)); d.f.push_back(rnd(0, 10));
    d.nk0.push_back(typed && rng() % 50 == 0); d.nk1.push_back(typed && rng() % 33 == 0);
    d.nv.push_back(typed && rng() % 6 == 0); d.nw.push_back(typed && rng() % 9 == 0); d.nf.push_back(typed && rng() % 20 == 0);
  }
  d.b_k0 = d.k0; d.b_k1 = d.k1; d.b_v = d.v; d.b_w = d.w; d.b_f = d.f;
  for (long long i = 0; i < n; i++) { d.t_k0.push_back((int32_t)d.k0[i]); d.t_k1.push_back((int16_t)d.k1[i]); d.t_w.push_back((int32_t)d.w[i]); d.t_f.push_back((int8_t)d.f[i]); }
  d.vb[0] = bitmap(d.nk0); d.vb[1] = bitmap(d.nk1); d.vb[2] = bitmap(d.nv); d.vb[3] = bitmap(d.nw); d.vb[4] = bitmap(d.nf);
  return d;
}

enum AccSpec { SUM_V, SUM_W, COUNT_V, COUNT_STAR };
struct Config {
  std::string name; int nkeys; bool typed; std::vector<AccSpec> accs; int nfilt; int form;   // form: 0 hash, 1 dense smem, 2 dense row, 3 dense gang (lean), 4 dense hot (lean)
  std::vector<int> dense_src;                                                            // dense entry words (-1 rows, -2 pad, j acc, 2+j valid counter)
  long long r0, r1;
};

struct Group { long long rows = 0; unsigned long long sum[2] = {0, 0}; long long nvalid[2] = {0, 0}; };
using Key = std::tuple<unsigned, long long, long long>;       // (key-is-NULL bits, k0, k1)

bool run(const Config& c, unsigned seed, bool via_dispatcher) {
  const long long n = 6000;
  Data d = make_data(n, c.r0, c.r1, c.typed, seed);
  // ---- columns: slots 0 k0, 1 k1, 2 v, 3 w, 4 f ----
  ColTable ct{};
  if (c.typed) {
    ct.col[0] = {d.t_k0.data(), d.vb[0].data(), 0, 0}; ct.col[1] = {d.t_k1.data(), d.vb[1].data(), 0, 0}; ct.col[2] = {d.b_v.data(), d.vb[2].data(), 0, 0};
    ct.col[3] = {d.t_w.data(), d.vb[3].data(), 0, 0}; ct.col[4] = {d.t_f.data(), d.vb[4].data(), 0, 0};
  } else {
    ct.col[0] = {d.b_k0.data(), nullptr, 0, 0}; ct.col[1] = {d.b_k1.data(), nullptr, 0, 0}; ct.col[2] = {d.b_v.data(), nullptr, 0, 0};
    ct.col[3] = {d.b_w.data(), nullptr, 0, 0}; ct.col[4] = {d.b_f.data(), nullptr, 0, 0};
  }
  const int nacc = (int)c.accs.size();
  FastSpec fs{}; fs.nkeys = c.nkeys; fs.nacc = nacc; fs.nfilt = c.nfilt; fs.lean = c.typed ? 0 : 1;
  fs.key_col[0] = 0; fs.key_phys[0] = c.typed ? PH_I32 : PH_I64; fs.key_col[1] = 1; fs.key_phys[1] = c.typed ? PH_I16 : PH_I64;
  AggLayout lay{}; lay.nkeys = c.nkeys; lay.nkw = c.nkeys; lay.nacc = nacc; lay.kstride = c.nkeys == 1 ? 2 : 4; lay.astride = 2;
  for (int j = 0; j < nacc; j++) {
    const AccSpec a = c.accs[j];
    fs.acc[j].kind = (a == SUM_V || a == SUM_W) ? FAST_ACC_ADD : FAST_ACC_COUNT;
    fs.acc[j].col = a == SUM_V || a == COUNT_V ? 2 : a == SUM_W ? 3 : -1;
    fs.acc[j].phys = a == SUM_W && c.typed ? PH_I32 : PH_I64;
    fs.acc[j].word = (uint8_t)j; fs.acc[j].vbit = fs.acc[j].kind == FAST_ACC_ADD ? (uint8_t)j : 0xFF;
    lay.acc[j].word = (uint8_t)j; lay.acc[j].vbit = fs.acc[j].vbit;
  }
  if (c.nfilt >= 1) { fs.filt[0].col = 4; fs.filt[0].phys = c.typed ? PH_I8 : PH_I64; fs.filt[0].op = CMP_GE; fs.filt[0].lit = 2; }
  if (c.nfilt >= 2) { fs.filt[1].col = 4; fs.filt[1].phys = c.typed ? PH_I8 : PH_I64; fs.filt[1].op = CMP_NE; fs.filt[1].lit = 7; }
  // merged per-column intervals, as stages.cu derives them: a `!=` term keeps the per-conjunct kernels (nfcol = -1)
  fs.nfcol = c.nfilt >= 2 ? -1 : c.nfilt;
  if (c.nfilt == 1) { fs.frange[0].col = 4; fs.frange[0].phys = fs.filt[0].phys; fs.frange[0].lo = 2; fs.frange[0].span = (unsigned long long)INT64_MAX - 2ull; }
  std::vector<unsigned long long> sink((size_t)FAST_SINK_WARPS * 4, 0); fs.sink = sink.data();
  // ---- hash table ----
  const uint64_t cap = 1 << 15;
  std::vector<unsigned long long> keys(cap * lay.kstride, 0), accs(cap * lay.astride, 0), counters(8, 0);
  std::vector<uint32_t> deferred((size_t)n, 0);
  AggTable tab{}; tab.keys = keys.data(); tab.accs = accs.data(); tab.capacity = cap; tab.max_groups = cap / 2; tab.counters = counters.data(); tab.deferred = deferred.data();
  // ---- dense table: range decided from the first two thirds of the rows (non-null keys), like decide_dense ----
  std::vector<unsigned long long> dtab;
  const int G = c.dense_src.empty() ? 0 : (c.dense_src.size() <= 2 ? 2 : 4);
  if (c.form != 0) {
    long long mn[2] = {INT64_MAX, INT64_MAX}, mx[2] = {INT64_MIN, INT64_MIN};
    for (long long i = 0; i < n * 2 / 3; i++) {
      if (!d.nk0[i]) { mn[0] = std::min(mn[0], d.k0[i]); mx[0] = std::max(mx[0], d.k0[i]); }
      if (!d.nk1[i]) { mn[1] = std::min(mn[1], d.k1[i]); mx[1] = std::max(mx[1], d.k1[i]); }
    }
    fs.dense = 1; fs.dense_stride = (int8_t)G;
    fs.dense_base = mn[0] - 2; fs.dense_cap0 = (unsigned long long)(mx[0] - mn[0] + 5);
    fs.dense_base1 = c.nkeys == 2 ? mn[1] - 1 : 0; fs.dense_r1 = c.nkeys == 2 ? (unsigned long long)(mx[1] - mn[1] + 3) : 1;
    fs.dense_cap = fs.dense_cap0 * fs.dense_r1;
    for (int m = 0; m < 4; m++) fs.dense_word_src[m] = m < (int)c.dense_src.size() ? (int8_t)c.dense_src[m] : (int8_t)-2;
    fs.dense_presence_word = 0;
    dtab.assign((size_t)fs.dense_cap * G, 0); fs.dense_tab = dtab.data();
    if (c.form == 1 && fs.dense_cap * G > (unsigned long long)DS_MAX_WORDS) { printf("  %s: table too large for the shared-memory form\n", c.name.c_str()); return false; }
    if (c.form == 4) fs.hot_cache = 1;
  }
  // ---- run ----
  if (via_dispatcher) {                                          // the product's own dispatcher picks the kernel form
    if (launch_agg_fast_update(ct, fs, lay, tab, 0, n, nullptr) != 1) { printf("  %s: dispatcher launched nothing\n", c.name.c_str()); return false; }
  } else {
  const unsigned grid = 3;
#define RUN(KERNEL, BLOCK) emu::launch(grid, BLOCK, [&] { KERNEL(ct, fs, lay, tab, 0, n); })
  const int nk = c.nkeys;
  if (c.form == 0) {
    if (c.typed) { if (nk == 1) { if (nacc == 2) RUN((agg_lean_hash_kernel<1, 2, true>), LH_BLOCK); else RUN((agg_lean_hash_kernel<1, 1, true>), LH_BLOCK); }
                   else { if (nacc == 2) RUN((agg_lean_hash_kernel<2, 2, true>), LH_BLOCK); else RUN((agg_lean_hash_kernel<2, 1, true>), LH_BLOCK); } }
    else { if (nk == 1) { if (nacc == 2) RUN((agg_lean_hash_kernel<1, 2, false>), LH_BLOCK); else RUN((agg_lean_hash_kernel<1, 1, false>), LH_BLOCK); }
           else { if (nacc == 2) RUN((agg_lean_hash_kernel<2, 2, false>), LH_BLOCK); else RUN((agg_lean_hash_kernel<2, 1, false>), LH_BLOCK); } }
  } else if (c.form == 1) {
    if (c.typed) { if (nk == 1) { if (nacc == 2) RUN((agg_dense_smem_kernel<2, true, 1>), FA_BLOCK); else RUN((agg_dense_smem_kernel<1, true, 1>), FA_BLOCK); }
                   else { if (nacc == 2) RUN((agg_dense_smem_kernel<2, true, 2>), FA_BLOCK); else RUN((agg_dense_smem_kernel<1, true, 2>), FA_BLOCK); } }
    else { if (nk == 1) { if (nacc == 2) RUN((agg_dense_smem_kernel<2, false, 1>), FA_BLOCK); else RUN((agg_dense_smem_kernel<1, false, 1>), FA_BLOCK); }
           else { if (nacc == 2) RUN((agg_dense_smem_kernel<2, false, 2>), FA_BLOCK); else RUN((agg_dense_smem_kernel<1, false, 2>), FA_BLOCK); } }
  } else if (c.form == 2) {
#define ROW(NACC, NK, GG) do { if (c.typed) RUN((agg_dense_row_kernel<NACC, NK, GG, true>), FA_BLOCK); else RUN((agg_dense_row_kernel<NACC, NK, GG, false>), FA_BLOCK); } while (0)
    if (nk == 1) { if (nacc == 2) { if (G == 2) ROW(2, 1, 2); else ROW(2, 1, 4); } else { if (G == 2) ROW(1, 1, 2); else ROW(1, 1, 4); } }
    else { if (nacc == 2) { if (G == 2) ROW(2, 2, 2); else ROW(2, 2, 4); } else { if (G == 2) ROW(1, 2, 2); else ROW(1, 2, 4); } }
  } else if (c.form == 3) {
    if (nacc == 2) { if (G == 2) RUN((agg_lean_dense_kernel<2, 2, 1>), FA_BLOCK); else RUN((agg_lean_dense_kernel<2, 4, 1>), FA_BLOCK); }
    else { if (G == 2) RUN((agg_lean_dense_kernel<1, 2, 1>), FA_BLOCK); else RUN((agg_lean_dense_kernel<1, 4, 1>), FA_BLOCK); }
  } else {
#define HOT(NACC, NK, GG) RUN((agg_dense_hot_kernel<NACC, NK, GG>), FA_BLOCK)
    if (nk == 1) { if (nacc == 2) { if (G == 2) HOT(2, 1, 2); else HOT(2, 1, 4); } else { if (G == 2) HOT(1, 1, 2); else HOT(1, 1, 4); } }
    else { if (nacc == 2) { if (G == 2) HOT(2, 2, 2); else HOT(2, 2, 4); } else { if (G == 2) HOT(1, 2, 2); else HOT(1, 2, 4); } }
  }
  }
  // ---- expected ----
  std::map<Key, Group> exp;
  for (long long i = 0; i < n; i++) {
    bool alive = true;
    if (c.nfilt >= 1) alive = alive && !d.nf[i] && d.f[i] >= 2;
    if (c.nfilt >= 2) alive = alive && !d.nf[i] && d.f[i] != 7;
    if (!alive) continue;
    const unsigned kn = (d.nk0[i] ? 1u : 0u) | ((c.nkeys == 2 && d.nk1[i]) ? 2u : 0u);
    Group& g = exp[Key(kn, d.nk0[i] ? 0 : d.k0[i], c.nkeys == 2 ? (d.nk1[i] ? 0 : d.k1[i]) : 0)];
    g.rows++;
    for (int j = 0; j < nacc; j++) {
      const AccSpec a = c.accs[j];
      const bool valid = a == COUNT_STAR ? true : (a == SUM_W ? !d.nw[i] : !d.nv[i]);
      if (!valid) continue;
      g.nvalid[j]++;
      g.sum[j] += a == SUM_V ? (unsigned long long)d.v[i] : a == SUM_W ? (unsigned long long)d.w[i] : 1ULL;
    }
  }
  // ---- actual: dense entries + hashed slots ----
  std::map<Key, Group> got; int errors = 0;
  auto complain = [&](const char* what, const Key& k) { if (errors++ < 5) printf("  %s: %s at key (null=%u, %lld, %lld)\n", c.name.c_str(), what, std::get<0>(k), std::get<1>(k), std::get<2>(k)); };
  if (c.form != 0) {
    for (unsigned long long e = 0; e < fs.dense_cap; e++) {
      const unsigned long long* w = &dtab[e * G];
      bool any = false; for (int m = 0; m < G; m++) any |= w[m] != 0;
      if (!any) continue;
      const Key k(0, fs.dense_base + (long long)(e / fs.dense_r1), c.nkeys == 2 ? fs.dense_base1 + (long long)(e % fs.dense_r1) : 0);
      auto it = exp.find(k);
      if (it == exp.end()) { complain("unexpected dense entry", k); continue; }
      const Group& g = it->second;
      for (int m = 0; m < G; m++) {
        const int src = fs.dense_word_src[m];
        const unsigned long long want = src == -1 ? (unsigned long long)g.rows : src == -2 ? 0ULL : src >= 2 ? (unsigned long long)g.nvalid[src - 2]
                                        : (fs.acc[src].kind == FAST_ACC_ADD ? g.sum[src] : (unsigned long long)g.nvalid[src]);
        if (w[m] != want) complain("wrong dense word", k);
      }
      got[k] = g;
    }
  }
  long long hashed = 0;
  for (uint64_t s = 0; s < cap; s++) {
    const unsigned long long hdr = keys[s * lay.kstride];
    if ((unsigned)hdr == 0) continue;
    if (!((unsigned)hdr & 0x80000000u)) { printf("  %s: slot %llu left locked\n", c.name.c_str(), (unsigned long long)s); errors++; continue; }
    hashed++;
    const unsigned flags = (unsigned)(hdr >> 32);
    const Key k(flags >> 16, (long long)keys[s * lay.kstride + 1], c.nkeys == 2 ? (long long)keys[s * lay.kstride + 2] : 0);
    auto it = exp.find(k);
    if (it == exp.end()) { complain("unexpected hashed group", k); continue; }
    if (got.count(k)) { complain("group both dense and hashed / hashed twice", k); continue; }
    const Group& g = it->second;
    for (int j = 0; j < nacc; j++) {
      const unsigned long long want = fs.acc[j].kind == FAST_ACC_ADD ? g.sum[j] : (unsigned long long)g.nvalid[j];
      if (accs[s * lay.astride + j] != want) complain("wrong hashed accumulator", k);
      if (fs.acc[j].vbit != 0xFF && (((flags >> fs.acc[j].vbit) & 1) != (g.nvalid[j] > 0))) complain("wrong accumulator-valid bit", k);
    }
    got[k] = g;
  }
  if (counters[0] != (unsigned long long)hashed) { printf("  %s: group counter %llu != %lld hashed slots\n", c.name.c_str(), counters[0], hashed); errors++; }
  if (counters[1] != 0) { printf("  %s: %llu rows deferred (table was sized for all groups)\n", c.name.c_str(), counters[1]); errors++; }
  for (auto& kv : exp) if (!got.count(kv.first)) complain("missing group", kv.first);
  printf("%-58s %-10s %s  (%zu groups, %lld hashed)\n", c.name.c_str(), via_dispatcher ? "dispatcher" : "forced", errors ? "FAIL" : "ok", exp.size(), hashed);
  return errors == 0;
}

}  // namespace

int main(int argc, char** argv) {
  const std::string only = argc > 1 ? argv[1] : "";
  std::vector<Config> cs;
  const long long TINY0 = 12, TINY1 = 3, MID0 = 700, MID1 = 6;
  for (int nk = 1; nk <= 2; nk++)
    for (int typed = 0; typed <= 1; typed++)
      for (int nf = 0; nf <= 2; nf++) {
        const std::string tag = std::string(nk == 1 ? "1key" : "2keys") + (typed ? " typed" : " lean") + (nf == 2 ? " filt" : nf == 1 ? " filt1" : "");
        cs.push_back({"hash   sum(v),count(v)        " + tag, nk, (bool)typed, {SUM_V, COUNT_V}, nf, 0, {}, typed ? 1LL << 30 : 1LL << 40, 9});     // typed keys are stored as int32
        cs.push_back({"hash   sum(v)                 " + tag, nk, (bool)typed, {SUM_V}, nf, 0, {}, 5000, 9});
        cs.push_back({"smem   {sum,count*}           " + tag, nk, (bool)typed, {SUM_V, COUNT_STAR}, nf, 1, {0, 1}, TINY0, TINY1});
        cs.push_back({"smem   {rows,sum,cnt(v),pad}  " + tag, nk, (bool)typed, {SUM_V, COUNT_V}, nf, 1, {-1, 0, 1, -2}, TINY0, TINY1});
        cs.push_back({"smem   {rows,sum,nvalid,pad}  " + tag, nk, (bool)typed, {SUM_V}, nf, 1, {-1, 0, 2, -2}, TINY0, TINY1});
        cs.push_back({"row    {sum,count*}           " + tag, nk, (bool)typed, {SUM_V, COUNT_STAR}, nf, 2, {0, 1}, MID0, MID1});
        cs.push_back({"row    {rows,sum}             " + tag, nk, (bool)typed, {SUM_V}, nf, 2, {-1, 0}, MID0, MID1});
        cs.push_back({"row    {rows,sum,cnt(v),pad}  " + tag, nk, (bool)typed, {SUM_V, COUNT_V}, nf, 2, {-1, 0, 1, -2}, MID0, MID1});
        cs.push_back({"row    {rows,sum,sum(w),nval0}" + tag, nk, (bool)typed, {SUM_V, SUM_W}, nf, 2, {-1, 0, 1, 2}, MID0, MID1});
        cs.push_back({"row    {rows,sum,nvalid,pad}  " + tag, nk, (bool)typed, {SUM_V}, nf, 2, {-1, 0, 2, -2}, MID0, MID1});
        if (!typed) {
          cs.push_back({"hot    {sum,count*}           " + tag, nk, false, {SUM_V, COUNT_STAR}, nf, 4, {0, 1}, MID0, MID1});
          cs.push_back({"hot    {rows,sum,cnt(v),pad}  " + tag, nk, false, {SUM_V, COUNT_V}, nf, 4, {-1, 0, 1, -2}, MID0, MID1});
          cs.push_back({"hot    {rows,sum,nvalid,pad}  " + tag, nk, false, {SUM_V}, nf, 4, {-1, 0, 2, -2}, MID0, MID1});
          if (nk == 1) {
            cs.push_back({"gang   {sum,count*}           " + tag, 1, false, {SUM_V, COUNT_STAR}, nf, 3, {0, 1}, MID0, MID1});
            cs.push_back({"gang   {rows,sum,cnt(v),pad}  " + tag, 1, false, {SUM_V, COUNT_V}, nf, 3, {-1, 0, 1, -2}, MID0, MID1});
            cs.push_back({"gang   {rows,sum,nvalid,pad}  " + tag, 1, false, {SUM_V}, nf, 3, {-1, 0, 2, -2}, MID0, MID1});
          }
        }
      }
  int failed = 0, ran = 0;
  for (size_t i = 0; i < cs.size(); i++) {
    if (!only.empty() && cs[i].name.find(only) == std::string::npos) continue;
    ran += 2;
    if (!run(cs[i], 1000 + (unsigned)i, false)) failed++;
    if (!run(cs[i], 1000 + (unsigned)i, true)) failed++;
  }
  // the key-range and skew-probe launchers (decide_dense's inputs)
  if (only.empty()) {
    ran++;
    Data d = make_data(50000, 700, 6, true, 7);
    DevCol kc{d.t_k0.data(), d.vb[0].data(), 0, 0};
    long long out[3] = {INT64_MAX, INT64_MIN, 0};
    launch_key_range(kc, PH_I32, d.n, out, nullptr);
    long long mn = INT64_MAX, mx = INT64_MIN, cnt = 0;
    for (long long i = 0; i < d.n; i++) if (!d.nk0[i]) { mn = std::min(mn, d.k0[i]); mx = std::max(mx, d.k0[i]); cnt++; }
    std::vector<unsigned> hist(65536 + 1, 0);
    DevCol kcs[2] = {kc, kc}; const uint8_t ph[2] = {PH_I32, PH_I32};
    launch_key_skew_probe(kcs, ph, 1, d.n, hist.data(), nullptr);
    std::map<long long, unsigned> freq; for (long long i = 0; i < d.n; i++) if (!d.nk0[i]) freq[d.k0[i]]++;
    unsigned top = 0; for (auto& kv : freq) top = std::max(top, kv.second);
    unsigned long long total = 0; for (int i = 0; i < 65536; i++) total += hist[i];
    const bool ok = out[0] == mn && out[1] == mx && out[2] == cnt && total == (unsigned long long)cnt && hist[65536] >= top && hist[65536] <= 4 * top + 16;
    printf("%-58s %-10s %s\n", "key range + skew probe launchers", "", ok ? "ok" : "FAIL");
    failed += !ok;
  }
  printf("%d configurations, %d failed\n", ran, failed);
  return failed ? 1 : 0;
}
