/*
 * CPU ORACLE (C) — TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY.  Never linked into the product.
 *
 * A plain-C restatement of the reference's CPU *algorithm* for the hot path (kwai/blaze = Apache
 * Auron @ d1eaef148a58), structured like the Rust code so that its timing is a fair "port" baseline
 * (bench.py cpu_baseline.kind = "port"; the reference binary itself cannot be built here: no
 * cargo/rustc):
 *
 *   HashAgg   10,000-row batches (datafusion-ext-commons/src/lib.rs:74-77);
 *             row-encoded keys: non-null int64 -> 0x01 || big-endian with the sign bit flipped
 *             (arrow-row, agg_ctx.rs:219-231); keys kept inline in 32-byte cells (agg_table.rs:61-62);
 *             open addressing over 64-byte groups {8 x u32 hash, 8 x u32 record id}, capacity =
 *             next_pow2(max(n,128)*2/8) groups => load <= 0.5, hash top bit forced, 0 = empty,
 *             first empty lane inserts, else next group, software prefetch 4 ahead
 *             (agg_hash_map.rs:66-136,228-234); dense record ids in insertion order;
 *             columnar accumulators: i64 sum + validity bit, i64 count
 *             (acc.rs:243-280, sum.rs:90-115, count.rs:90-126).
 *   Filter+Project  separate passes per batch: predicate mask -> null->false -> compact each column ->
 *             project (cached_exprs_evaluator.rs:90-166,495-524).
 *   T threads = one reference "task" per thread over N/T rows (Spark task parallelism, rt.rs:110-130): Partial per
 *             task, partial records bucketed by key hash (the shuffle), then one Final-mode merge per reduce
 *             partition, also T threads (agg_ctx.rs:276-301).
 *
 * The slot/bucket hash (foldhash 0.1.5 in the reference) is not observable in results; a folded
 * multiply of the same flavour is used.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BATCH_SIZE 10000
#define KEY_CELL 32          /* SmallVec<u8,24> is 32 bytes (acc.rs:79-80) */

typedef struct {
  uint32_t* groups;          /* ngroups * 16 u32: [8 hashes][8 record ids] */
  uint64_t ngroups_mask;
  uint64_t ngroups;
  uint8_t* keys;             /* nrec * KEY_CELL: [len u8][bytes...] */
  int64_t* sums; uint8_t* sum_valid; int64_t* counts;
  uint64_t nrec, cap_rec;
} table_t;

static inline uint64_t fold_mul(uint64_t a, uint64_t b) { __uint128_t m = (__uint128_t)a * b; return (uint64_t)m ^ (uint64_t)(m >> 64); }
static inline uint32_t key_hash(const uint8_t* k, int len) {
  uint64_t a = 0, b = 0;
  memcpy(&a, k, len < 8 ? len : 8);
  if (len > 8) memcpy(&b, k + len - 8, 8);
  uint64_t h = fold_mul(a ^ 0x3F6F1B93243F6A88ULL, b ^ 0x13198A2E03707344ULL ^ (uint64_t)len);
  return (uint32_t)h | 0x80000000u;   /* top bit forced, 0 = empty (agg_hash_map.rs:228-234) */
}

static void table_alloc_groups(table_t* t, uint64_t n_records) {
  uint64_t want = (n_records < 128 ? 128 : n_records) * 2 / 8, g = 1;
  while (g < want) g <<= 1;
  t->ngroups = g; t->ngroups_mask = g - 1;
  if (posix_memalign((void**)&t->groups, 64, g * 64)) abort();
  memset(t->groups, 0, g * 64);
}
static void table_init(table_t* t) {
  memset(t, 0, sizeof(*t));
  table_alloc_groups(t, 128);
  t->cap_rec = 1024;
  t->keys = malloc(t->cap_rec * KEY_CELL); t->sums = malloc(t->cap_rec * 8); t->counts = malloc(t->cap_rec * 8); t->sum_valid = malloc(t->cap_rec);
}
static void table_free(table_t* t) { free(t->groups); free(t->keys); free(t->sums); free(t->counts); free(t->sum_valid); }

static inline void insert_hash(table_t* t, uint32_t h, uint32_t rec) {
  uint64_t g = h & t->ngroups_mask;
  for (;;) {
    uint32_t* grp = t->groups + g * 16;
    for (int l = 0; l < 8; l++) if (grp[l] == 0) { grp[l] = h; grp[8 + l] = rec; return; }
    g = (g + 1) & t->ngroups_mask;
  }
}
static void table_reserve(table_t* t, uint64_t extra) {            /* reserve + rehash (agg_hash_map.rs:66-75,138-169) */
  uint64_t need = t->nrec + extra;
  if (need * 2 / 8 > t->ngroups) {
    free(t->groups);
    table_alloc_groups(t, need);
    for (uint64_t r = 0; r < t->nrec; r++) { const uint8_t* k = t->keys + r * KEY_CELL; insert_hash(t, key_hash(k + 1, k[0]), (uint32_t)r); }
  }
  if (need > t->cap_rec) {
    while (t->cap_rec < need) t->cap_rec *= 2;
    t->keys = realloc(t->keys, t->cap_rec * KEY_CELL); t->sums = realloc(t->sums, t->cap_rec * 8);
    t->counts = realloc(t->counts, t->cap_rec * 8); t->sum_valid = realloc(t->sum_valid, t->cap_rec);
  }
}

/* upsert_one_impl (agg_hash_map.rs:102-136): 8-lane hash compare, memcmp confirm, first empty lane inserts */
static inline uint32_t upsert_one(table_t* t, const uint8_t* key, int len, uint32_t h) {
  uint64_t g = h & t->ngroups_mask;
  for (;;) {
    uint32_t* grp = t->groups + g * 16;
    uint32_t eq = 0, empty = 0;
    for (int l = 0; l < 8; l++) { eq |= (uint32_t)(grp[l] == h) << l; empty |= (uint32_t)(grp[l] == 0) << l; }   /* Simd<u32,8> */
    while (eq) {
      int l = __builtin_ctz(eq); eq &= eq - 1;
      const uint8_t* k = t->keys + (uint64_t)grp[8 + l] * KEY_CELL;
      if (k[0] == len && memcmp(k + 1, key, len) == 0) return grp[8 + l];
    }
    if (empty) {
      int l = __builtin_ctz(empty);
      uint32_t rec = (uint32_t)t->nrec++;
      grp[l] = h; grp[8 + l] = rec;
      uint8_t* k = t->keys + (uint64_t)rec * KEY_CELL; k[0] = (uint8_t)len; memcpy(k + 1, key, len);
      t->sums[rec] = 0; t->sum_valid[rec] = 0; t->counts[rec] = 0;       /* AccColumn::resize defaults */
      return rec;
    }
    g = (g + 1) & t->ngroups_mask;
  }
}

/* arrow-row encoding of one nullable int64 key: NULL -> 0x00 + 8 zero bytes, else 0x01 + BE(sign-flipped) */
static inline void encode_key(int64_t v, int valid, uint8_t* out) {
  if (!valid) { memset(out, 0, 9); return; }
  uint64_t u = (uint64_t)v ^ 0x8000000000000000ULL;
  out[0] = 1;
  for (int i = 0; i < 8; i++) out[1 + i] = (uint8_t)(u >> (56 - 8 * i));
}
static inline int64_t decode_key(const uint8_t* k, int* valid) {
  *valid = k[0] != 0;
  uint64_t u = 0; for (int i = 0; i < 8; i++) u = (u << 8) | k[1 + i];
  return (int64_t)(u ^ 0x8000000000000000ULL);
}
static inline int bit_get(const uint8_t* bits, int64_t i) { return bits ? (bits[i >> 3] >> (i & 7)) & 1 : 1; }

/* HashingData::update_batch for one <=10,000-row batch: SUM(v), COUNT(v) GROUP BY k */
static void update_batch(table_t* t, const int64_t* k, const uint8_t* kvalid, const int64_t* v, const uint8_t* vvalid, int64_t base, int n,
                         uint8_t* rows /* n*9 */, uint32_t* hashes, uint32_t* recs) {
  for (int i = 0; i < n; i++) { encode_key(k[base + i], bit_get(kvalid, base + i), rows + 9 * i); hashes[i] = key_hash(rows + 9 * i, 9); }   /* K4 */
  table_reserve(t, (uint64_t)n);
  for (int i = 0; i < n; i++) {                                                                                                        /* K5 */
    if (i + 4 < n) __builtin_prefetch(t->groups + (hashes[i + 4] & t->ngroups_mask) * 16, 1);                                           /* :79,93-98 */
    recs[i] = upsert_one(t, rows + 9 * i, 9, hashes[i]);
  }
  for (int i = 0; i < n; i++) {                                                                                                        /* K6: AggSum::partial_update */
    if (bit_get(vvalid, base + i)) {
      uint32_t r = recs[i];
      if (t->sum_valid[r]) t->sums[r] = (int64_t)((uint64_t)t->sums[r] + (uint64_t)v[base + i]); else { t->sums[r] = v[base + i]; t->sum_valid[r] = 1; }
    }
  }
  for (int i = 0; i < n; i++) t->counts[recs[i]] += bit_get(vvalid, base + i);                                                          /* AggCount::partial_update */
}

typedef struct task_s {
  const int64_t* k; const uint8_t* kvalid; const int64_t* v; const uint8_t* vvalid; int64_t begin, end; table_t t;
  /* shuffle write: records of this task bucketed by the final partition that owns their key */
  int nparts; uint32_t* bucket_off; uint32_t* bucket_idx;
  /* final stage (one reduce partition per thread) */
  struct task_s* all; int self; table_t fin;
} task_t;

static void* task_main(void* arg) {
  task_t* ta = (task_t*)arg;
  table_init(&ta->t);
  uint8_t* rows = malloc(BATCH_SIZE * 9); uint32_t* hashes = malloc(BATCH_SIZE * 4); uint32_t* recs = malloc(BATCH_SIZE * 4);
  for (int64_t b = ta->begin; b < ta->end; b += BATCH_SIZE) {
    int n = (int)(ta->end - b < BATCH_SIZE ? ta->end - b : BATCH_SIZE);
    update_batch(&ta->t, ta->k, ta->kvalid, ta->v, ta->vvalid, b, n, rows, hashes, recs);
  }
  free(rows); free(hashes); free(recs);
  /* "shuffle write": partition the partial records by owner = hash % nparts (Spark: murmur3 pmod, shuffle/mod.rs:163-188) */
  int P = ta->nparts; table_t* t = &ta->t;
  ta->bucket_off = calloc((size_t)P + 1, 4); ta->bucket_idx = malloc((t->nrec ? t->nrec : 1) * 4);
  for (uint64_t r = 0; r < t->nrec; r++) { const uint8_t* k = t->keys + r * KEY_CELL; ta->bucket_off[1 + (key_hash(k + 1, k[0]) >> 3) % (uint32_t)P]++; }
  for (int p = 0; p < P; p++) ta->bucket_off[p + 1] += ta->bucket_off[p];
  uint32_t* cur = malloc((size_t)P * 4); memcpy(cur, ta->bucket_off, (size_t)P * 4);
  for (uint64_t r = 0; r < t->nrec; r++) { const uint8_t* k = t->keys + r * KEY_CELL; ta->bucket_idx[cur[(key_hash(k + 1, k[0]) >> 3) % (uint32_t)P]++] = (uint32_t)r; }
  free(cur);
  return NULL;
}

/* Final-mode AggExec of one reduce partition: merge the partial records every task bucketed for it (agg_ctx.rs:276-301) */
static void* final_main(void* arg) {
  task_t* me = (task_t*)arg; table_t* fin = &me->fin;
  table_init(fin);
  for (int i = 0; i < me->nparts; i++) {
    task_t* src = &me->all[i]; table_t* p = &src->t;
    uint32_t lo = src->bucket_off[me->self], hi = src->bucket_off[me->self + 1];
    table_reserve(fin, hi - lo);
    for (uint32_t x = lo; x < hi; x++) {
      uint64_t r = src->bucket_idx[x];
      const uint8_t* key = p->keys + r * KEY_CELL;
      uint32_t rec = upsert_one(fin, key + 1, key[0], key_hash(key + 1, key[0]));
      if (p->sum_valid[r]) { if (fin->sum_valid[rec]) fin->sums[rec] = (int64_t)((uint64_t)fin->sums[rec] + (uint64_t)p->sums[r]); else { fin->sums[rec] = p->sums[r]; fin->sum_valid[rec] = 1; } }
      fin->counts[rec] += p->counts[r];
    }
  }
  return NULL;
}

/* Partial per task -> shuffle by key hash -> Final per partition, T threads in both stages (Spark's two-stage plan,
 * NativeAggBase.scala:129-135).  Outputs (caller-allocated, capacity >= number of distinct keys, or NULL to only
 * count): returns the number of groups.  key_valid/sum_valid are one byte per group. */
int64_t cpu_ref_hashagg_sum_count(const int64_t* k, const uint8_t* kvalid, const int64_t* v, const uint8_t* vvalid, int64_t n, int nthreads,
                                  int64_t* out_k, uint8_t* out_kvalid, int64_t* out_sum, uint8_t* out_sumvalid, int64_t* out_cnt, int64_t out_cap) {
  if (nthreads < 1) nthreads = 1;
  task_t* tasks = calloc((size_t)nthreads, sizeof(task_t));
  pthread_t* th = calloc((size_t)nthreads, sizeof(pthread_t));
  int64_t per = ((n + nthreads - 1) / nthreads + BATCH_SIZE - 1) / BATCH_SIZE * BATCH_SIZE;
  for (int i = 0; i < nthreads; i++) {
    tasks[i].k = k; tasks[i].kvalid = kvalid; tasks[i].v = v; tasks[i].vvalid = vvalid; tasks[i].nparts = nthreads; tasks[i].all = tasks; tasks[i].self = i;
    tasks[i].begin = per * i < n ? per * i : n; tasks[i].end = per * (i + 1) < n ? per * (i + 1) : n;
    if (nthreads == 1) task_main(&tasks[i]); else pthread_create(&th[i], NULL, task_main, &tasks[i]);
  }
  if (nthreads > 1) for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  for (int i = 0; i < nthreads; i++) { if (nthreads == 1) final_main(&tasks[i]); else pthread_create(&th[i], NULL, final_main, &tasks[i]); }
  if (nthreads > 1) for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  int64_t g = 0;
  for (int i = 0; i < nthreads; i++) g += (int64_t)tasks[i].fin.nrec;
  if (out_k && g <= out_cap) {
    int64_t o = 0;
    for (int i = 0; i < nthreads; i++) {
      table_t* fin = &tasks[i].fin;
      for (uint64_t r = 0; r < fin->nrec; r++, o++) {
        int valid; out_k[o] = decode_key(fin->keys + r * KEY_CELL + 1, &valid);
        out_kvalid[o] = (uint8_t)valid; out_sum[o] = fin->sum_valid[r] ? fin->sums[r] : 0; out_sumvalid[o] = fin->sum_valid[r]; out_cnt[o] = fin->counts[r];
      }
    }
  }
  for (int i = 0; i < nthreads; i++) { table_free(&tasks[i].t); table_free(&tasks[i].fin); free(tasks[i].bucket_off); free(tasks[i].bucket_idx); }
  free(tasks); free(th);
  return g;
}

/* FilterExec[a < thr] -> ProjectExec[a, a + b] over 10,000-row batches; separate mask / compact / project passes.
 * Returns the number of output rows; out_a/out_c have capacity n; validity as bytes (or NULL in / out). */
typedef struct { const int64_t* a; const uint8_t* av; const int64_t* b; const uint8_t* bv; int64_t thr; int64_t begin, end; int64_t* out_a; int64_t* out_c; uint8_t* out_cv; int64_t count; } fp_task_t;

static void* fp_main(void* arg) {
  fp_task_t* t = (fp_task_t*)arg;
  uint8_t* mask = malloc(BATCH_SIZE); int64_t* fa = malloc(BATCH_SIZE * 8); int64_t* fb = malloc(BATCH_SIZE * 8); uint8_t* fbv = malloc(BATCH_SIZE);
  int64_t out = t->begin;                                    /* each task writes its own disjoint output range, compacted later by the caller */
  for (int64_t base = t->begin; base < t->end; base += BATCH_SIZE) {
    int n = (int)(t->end - base < BATCH_SIZE ? t->end - base : BATCH_SIZE);
    for (int i = 0; i < n; i++) mask[i] = (uint8_t)(bit_get(t->av, base + i) && t->a[base + i] < t->thr);     /* filter_one_pred: null -> false */
    int m = 0;
    for (int i = 0; i < n; i++) if (mask[i]) fa[m++] = t->a[base + i];                                          /* filter_record_batch, column a */
    m = 0;
    for (int i = 0; i < n; i++) if (mask[i]) { fb[m] = t->b[base + i]; fbv[m] = (uint8_t)bit_get(t->bv, base + i); m++; }   /* column b */
    for (int i = 0; i < m; i++) t->out_a[out + i] = fa[i];                                                       /* projection a */
    for (int i = 0; i < m; i++) { t->out_c[out + i] = (int64_t)((uint64_t)fa[i] + (uint64_t)fb[i]); if (t->out_cv) t->out_cv[out + i] = fbv[i]; }   /* a + b (wrapping) */
    out += m;
  }
  t->count = out - t->begin;
  free(mask); free(fa); free(fb); free(fbv);
  return NULL;
}

int64_t cpu_ref_filter_project(const int64_t* a, const uint8_t* av, const int64_t* b, const uint8_t* bv, int64_t n, int64_t thr, int nthreads,
                               int64_t* out_a, int64_t* out_c, uint8_t* out_cv) {
  if (nthreads < 1) nthreads = 1;
  fp_task_t* tasks = calloc((size_t)nthreads, sizeof(fp_task_t)); pthread_t* th = calloc((size_t)nthreads, sizeof(pthread_t));
  int64_t per = ((n + nthreads - 1) / nthreads + BATCH_SIZE - 1) / BATCH_SIZE * BATCH_SIZE;
  for (int i = 0; i < nthreads; i++) {
    fp_task_t* t = &tasks[i];
    t->a = a; t->av = av; t->b = b; t->bv = bv; t->thr = thr; t->out_a = out_a; t->out_c = out_c; t->out_cv = out_cv;
    t->begin = per * i < n ? per * i : n; t->end = per * (i + 1) < n ? per * (i + 1) : n;
    if (nthreads == 1) fp_main(t); else pthread_create(&th[i], NULL, fp_main, t);
  }
  if (nthreads > 1) for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  int64_t total = tasks[0].count;                              /* stitch the per-task ranges together (order preserved) */
  for (int i = 1; i < nthreads; i++) {
    memmove(out_a + total, out_a + tasks[i].begin, (size_t)tasks[i].count * 8);
    memmove(out_c + total, out_c + tasks[i].begin, (size_t)tasks[i].count * 8);
    if (out_cv) memmove(out_cv + total, out_cv + tasks[i].begin, (size_t)tasks[i].count);
    total += tasks[i].count;
  }
  free(tasks); free(th);
  return total;
}


/* ---------------------------------------------------------------------------------------------------------------
 * M2 (TPC-DS q1 shape): FilterExec[f >= lo, f <= hi] -> AggExec SUM(v) GROUP BY k1, k2.
 * Per 10,000-row batch, as the reference operators do it:
 *   FilterExec: conjunct 1 over the batch -> mask; conjunct 2 evaluated on the selected rows only and scattered back
 *   (filter_one_pred / evaluate_selection, cached_exprs_evaluator.rs:495-524); NULL -> false; then the columns the
 *   parent needs (k1, k2, v — column pruning, column_pruning.rs:68-90) are compacted (filter_record_batch);
 *   AggExec: two int64 keys row-encoded to 18 bytes (arrow-row), 8-lane hash groups, SUM accumulator (sum.rs:90-115).
 * Then the same Partial per task -> bucket by key hash -> Final per partition structure as above.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct q1_task_s {
  const int64_t *f, *k1, *k2, *v; int64_t lo, hi, begin, end; table_t t;
  int nparts; uint32_t* bucket_off; uint32_t* bucket_idx; struct q1_task_s* all; int self; table_t fin;
} q1_task_t;

static void* q1_task_main(void* arg) {
  q1_task_t* ta = (q1_task_t*)arg; table_t* t = &ta->t;
  table_init(t);
  uint8_t* mask = malloc(BATCH_SIZE); int32_t* sel = malloc(BATCH_SIZE * 4);
  int64_t* c1 = malloc(BATCH_SIZE * 8); int64_t* c2 = malloc(BATCH_SIZE * 8); int64_t* cv = malloc(BATCH_SIZE * 8);
  uint8_t* rows = malloc(BATCH_SIZE * 18); uint32_t* hashes = malloc(BATCH_SIZE * 4); uint32_t* recs = malloc(BATCH_SIZE * 4);
  for (int64_t base = ta->begin; base < ta->end; base += BATCH_SIZE) {
    const int n = (int)(ta->end - base < BATCH_SIZE ? ta->end - base : BATCH_SIZE);
    for (int i = 0; i < n; i++) mask[i] = (uint8_t)(ta->f[base + i] >= ta->lo);                      /* conjunct 1 */
    int ns = 0; for (int i = 0; i < n; i++) if (mask[i]) sel[ns++] = i;                               /* selection of conjunct 2 */
    for (int s = 0; s < ns; s++) mask[sel[s]] = (uint8_t)(ta->f[base + sel[s]] <= ta->hi);            /* evaluate_selection + scatter */
    int m = 0;
    for (int i = 0; i < n; i++) if (mask[i]) c1[m++] = ta->k1[base + i];                              /* filter_record_batch, one pass per column */
    m = 0; for (int i = 0; i < n; i++) if (mask[i]) c2[m++] = ta->k2[base + i];
    m = 0; for (int i = 0; i < n; i++) if (mask[i]) cv[m++] = ta->v[base + i];
    if (m == 0) continue;                                                                             /* empty batches are dropped (execution_context.rs:713-716) */
    for (int i = 0; i < m; i++) { encode_key(c1[i], 1, rows + 18 * i); encode_key(c2[i], 1, rows + 18 * i + 9); hashes[i] = key_hash(rows + 18 * i, 18); }
    table_reserve(t, (uint64_t)m);
    for (int i = 0; i < m; i++) {
      if (i + 4 < m) __builtin_prefetch(t->groups + (hashes[i + 4] & t->ngroups_mask) * 16, 1);
      recs[i] = upsert_one(t, rows + 18 * i, 18, hashes[i]);
    }
    for (int i = 0; i < m; i++) {
      const uint32_t r = recs[i];
      if (t->sum_valid[r]) t->sums[r] = (int64_t)((uint64_t)t->sums[r] + (uint64_t)cv[i]); else { t->sums[r] = cv[i]; t->sum_valid[r] = 1; }
    }
  }
  free(mask); free(sel); free(c1); free(c2); free(cv); free(rows); free(hashes); free(recs);
  const int P = ta->nparts;
  ta->bucket_off = calloc((size_t)P + 1, 4); ta->bucket_idx = malloc((t->nrec ? t->nrec : 1) * 4);
  for (uint64_t r = 0; r < t->nrec; r++) { const uint8_t* k = t->keys + r * KEY_CELL; ta->bucket_off[1 + (key_hash(k + 1, k[0]) >> 3) % (uint32_t)P]++; }
  for (int p = 0; p < P; p++) ta->bucket_off[p + 1] += ta->bucket_off[p];
  uint32_t* cur = malloc((size_t)P * 4); memcpy(cur, ta->bucket_off, (size_t)P * 4);
  for (uint64_t r = 0; r < t->nrec; r++) { const uint8_t* k = t->keys + r * KEY_CELL; ta->bucket_idx[cur[(key_hash(k + 1, k[0]) >> 3) % (uint32_t)P]++] = (uint32_t)r; }
  free(cur);
  return NULL;
}

static void* q1_final_main(void* arg) {
  q1_task_t* me = (q1_task_t*)arg; table_t* fin = &me->fin;
  table_init(fin);
  for (int i = 0; i < me->nparts; i++) {
    q1_task_t* src = &me->all[i]; table_t* p = &src->t;
    const uint32_t lo = src->bucket_off[me->self], hi = src->bucket_off[me->self + 1];
    table_reserve(fin, hi - lo);
    for (uint32_t x = lo; x < hi; x++) {
      const uint64_t r = src->bucket_idx[x];
      const uint8_t* key = p->keys + r * KEY_CELL;
      const uint32_t rec = upsert_one(fin, key + 1, key[0], key_hash(key + 1, key[0]));
      if (p->sum_valid[r]) { if (fin->sum_valid[rec]) fin->sums[rec] = (int64_t)((uint64_t)fin->sums[rec] + (uint64_t)p->sums[r]); else { fin->sums[rec] = p->sums[r]; fin->sum_valid[rec] = 1; } }
    }
  }
  return NULL;
}

/* returns the number of groups; outputs (capacity out_cap, or NULL to only time): k1, k2, sum per group */
int64_t cpu_ref_q1_filter_agg(const int64_t* f, const int64_t* k1, const int64_t* k2, const int64_t* v, int64_t n, int64_t lo, int64_t hi, int nthreads,
                              int64_t* out_k1, int64_t* out_k2, int64_t* out_sum, int64_t out_cap) {
  if (nthreads < 1) nthreads = 1;
  q1_task_t* tasks = calloc((size_t)nthreads, sizeof(q1_task_t));
  pthread_t* th = calloc((size_t)nthreads, sizeof(pthread_t));
  const int64_t per = ((n + nthreads - 1) / nthreads + BATCH_SIZE - 1) / BATCH_SIZE * BATCH_SIZE;
  for (int i = 0; i < nthreads; i++) {
    q1_task_t* t = &tasks[i];
    t->f = f; t->k1 = k1; t->k2 = k2; t->v = v; t->lo = lo; t->hi = hi; t->nparts = nthreads; t->all = tasks; t->self = i;
    t->begin = per * i < n ? per * i : n; t->end = per * (i + 1) < n ? per * (i + 1) : n;
    if (nthreads == 1) q1_task_main(t); else pthread_create(&th[i], NULL, q1_task_main, t);
  }
  if (nthreads > 1) for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  for (int i = 0; i < nthreads; i++) { if (nthreads == 1) q1_final_main(&tasks[i]); else pthread_create(&th[i], NULL, q1_final_main, &tasks[i]); }
  if (nthreads > 1) for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  int64_t g = 0;
  for (int i = 0; i < nthreads; i++) g += (int64_t)tasks[i].fin.nrec;
  if (out_k1 && g <= out_cap) {
    int64_t o = 0;
    for (int i = 0; i < nthreads; i++) {
      table_t* fin = &tasks[i].fin;
      for (uint64_t r = 0; r < fin->nrec; r++, o++) {
        int valid; const uint8_t* key = fin->keys + r * KEY_CELL + 1;
        out_k1[o] = decode_key(key, &valid); out_k2[o] = decode_key(key + 9, &valid); out_sum[o] = fin->sums[r];
      }
    }
  }
  for (int i = 0; i < nthreads; i++) { table_free(&tasks[i].t); table_free(&tasks[i].fin); free(tasks[i].bucket_off); free(tasks[i].bucket_idx); }
  free(tasks); free(th);
  return g;
}
