"""Device-resident throughput of the other SURVEY §8d shapes (M0 filter+project, M2 q1-shaped fused
filter->agg, M1 through the hash path) — kernel-only numbers from b200q_metrics.hot_kernel_ns."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blaze_b200 import exprs as E, native, plans as PL, types as T

rows = int(os.environ.get("ROWS", 1 << 28))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(42)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0

REPS = int(os.environ.get("REPS", 0))      # override the repetitions of every shape (profiling)
ONLY = os.environ.get("SHAPES")       # comma-separated substrings of shape names to run

def run(name, plan_bytes, cols, alg_bytes_per_row, conf=None, reps=3, steady=False, valid=None):
    if ONLY and not any(t in name for t in ONLY.split(",")): return
    valid = valid or [None] * len(cols)
    reps = REPS or reps
    spec = [(c.data_ptr(), (vb.data_ptr() if vb is not None else 0), rows) for c, vb in zip(cols, valid)]
    keep = list(cols) + [vb for vb in valid if vb is not None]
    best = None; steady_ns = None
    for _ in range(reps):
        with native.NativeOp(plan_bytes, conf or native.default_conf(), 0) as op:
            op.push_device(native.DeviceBatch(spec, rows, 0, keepalive=keep))
            if steady:      # second pass over the same rows: every group already exists (steady-state cost, no inserts)
                op.sync(); m0 = op.metrics()
                op.push_device(native.DeviceBatch(spec, rows, 0, keepalive=keep))
                op.sync(); m1 = op.metrics()
                steady_ns = m1["hot_kernel_ns"] - m0["hot_kernel_ns"]
            op.finish()
            n_out = 0
            while True:
                o = op.pull_device()
                if o is None: break
                n_out += o.array.length; native.release_device_array(o)
            m = op.metrics()
        t = m["hot_kernel_ns"] / max(1, m["hot_kernel_launches"]) * m["hot_kernel_launches"]
        if best is None or t < best[0]: best = (t, m, n_out)
    t, m, n_out = best
    if steady:
        print(json.dumps({"shape": name + " [steady state, 2nd pass]", "rows": rows, "hot_kernel_ms": steady_ns / 1e6, "rows_per_s": rows / (steady_ns * 1e-9),
                          "alg_GBps": alg_bytes_per_row * rows / steady_ns, "frac_of_measured_hbm": alg_bytes_per_row * rows / steady_ns / peak}), flush=True)
        return
    gbs = alg_bytes_per_row * m["hot_kernel_rows"] / t
    print(json.dumps({"shape": name, "rows": rows, "out_rows": n_out, "hot_kernel_ms": t / 1e6, "rows_per_s": m["hot_kernel_rows"] / (t * 1e-9),
                      "alg_GBps": gbs, "frac_of_measured_hbm": gbs / peak, "fast_path_launches": m["fast_path_launches"], "launches": m["gpu_kernel_launches"]}), flush=True)

# M0: Filter[a < 500] -> Project[a, a + b]   (24 B/row at s = 0.5)
a = torch.randint(0, 1000, (rows,), dtype=torch.int64, device=dev, generator=g)
b = torch.randint(-2**31, 2**31, (rows,), dtype=torch.int64, device=dev, generator=g)
s0 = T.Schema([T.Field("a", T.int64, False), T.Field("b", T.int64, False)])
A, B = E.Column("a"), E.Column("b")
m0 = PL.ProjectExec([(A, "a"), (E.BinaryExpr(A, "Plus", B), "c")], PL.FilterExec([E.BinaryExpr(A, "Lt", E.Literal(500, T.int64))], PL.MemoryExec(s0)))
run("M0 filter+project s=0.5", m0.plan_bytes(), [a, b], 24.0)
del a, b
# M1 via hash path (dense disabled) and via the generic VM kernel
k = torch.randint(0, 1 << 20, (rows,), dtype=torch.int64, device=dev, generator=g)
v = torch.randint(-10**6, 10**6, (rows,), dtype=torch.int64, device=dev, generator=g)
s1 = T.Schema([T.Field("k", T.int64, False), T.Field("v", T.int64, False)])
aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], s1, T.int64)), E.AggExpr("c", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Column("v")], s1, T.int64))]
m1 = PL.AggExec(PL.HashAgg, [E.GroupingExpr("k", E.Column("k"))], aggs, True, PL.MemoryExec(s1))
run("M1 dense (lean, global table)", m1.plan_bytes(), [k, v], 16.0, native.default_conf(agg_initial_groups=1 << 20))
run("M1 hash (lean, compacted probes; first pass incl. 1M inserts)", m1.plan_bytes(), [k, v], 16.0, native.default_conf(agg_initial_groups=1 << 20, agg_dense_keys=0))
run("M1 hash", m1.plan_bytes(), [k, v], 16.0, native.default_conf(agg_initial_groups=1 << 20, agg_dense_keys=0), reps=1, steady=True)
for ig in (1 << 21, 1 << 22):
    run("M1 hash initial_groups=%d" % ig, m1.plan_bytes(), [k, v], 16.0, native.default_conf(agg_initial_groups=ig, agg_dense_keys=0), reps=1, steady=True)
# typed / nullable inputs through the hashed path: int32 key, 10 % NULL values (validity bitmaps)  (12.25 B/row)
k32 = k.to(torch.int32)
vbits = torch.full(((rows + 7) // 8,), 0xFF, dtype=torch.uint8, device=dev); vbits[::10] = 0
s1t = T.Schema([T.Field("k", T.int32, False), T.Field("v", T.int64, True)])
aggs_t = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], s1t, T.int64)), E.AggExpr("c", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Column("v")], s1t, T.int64))]
m1t = PL.AggExec(PL.HashAgg, [E.GroupingExpr("k", E.Column("k"))], aggs_t, True, PL.MemoryExec(s1t))
run("M1 typed hash (int32 key, nullable v)", m1t.plan_bytes(), [k32, v], 12.125, native.default_conf(agg_initial_groups=1 << 20, agg_dense_keys=0), reps=1, steady=True, valid=[None, vbits])
run("M1 typed dense (int32 key, nullable v)", m1t.plan_bytes(), [k32, v], 12.125, native.default_conf(agg_initial_groups=1 << 20), reps=2, valid=[None, vbits])
del k32, vbits
# low cardinality: 64 groups (every RED of a warp lands on a handful of sectors)
klow = torch.randint(0, 64, (rows,), dtype=torch.int64, device=dev, generator=g)
run("M1 low cardinality (64 groups) dense", m1.plan_bytes(), [klow, v], 16.0, native.default_conf())
run("M1 low cardinality (64 groups) hash", m1.plan_bytes(), [klow, v], 16.0, native.default_conf(agg_dense_keys=0), reps=1, steady=True)
del klow
# skewed keys: Zipf(1.1) ranks over 2^20 keys (continuous inverse-CDF approximation), rank r -> key (r * 2654435761) mod 2^20
u = torch.rand(rows, device=dev, generator=g, dtype=torch.float64)
nk, sz = float(1 << 20), 1.1
ranks = torch.clamp(((u * (nk ** (1 - sz) - 1) + 1) ** (1 / (1 - sz))).floor().to(torch.int64), 1, 1 << 20) - 1
kz = (ranks * 2654435761) % (1 << 20)
del u, ranks
run("M1 Zipf(1.1) keys, dense, hot-key cache off", m1.plan_bytes(), [kz, v], 16.0, native.default_conf(agg_initial_groups=1 << 20, agg_hot_key_cache=0), reps=1)
run("M1 Zipf(1.1) keys, dense + hot-key cache", m1.plan_bytes(), [kz, v], 16.0, native.default_conf(agg_initial_groups=1 << 20, agg_hot_key_cache=1), reps=1)
run("M1 Zipf(1.1) keys, hash", m1.plan_bytes(), [kz, v], 16.0, native.default_conf(agg_initial_groups=1 << 20, agg_dense_keys=0), reps=1, steady=True)
del kz
# the "fp64 SUM/AVG", "decimal128(17,2)" and MIN/MAX halves of the north_star target: wide tile kernels (kernels_tile.cu)
vf = v.to(torch.float64)
s1f = T.Schema([T.Field("k", T.int64, False), T.Field("v", T.float64, False)])
mkp = lambda sch, specs: PL.AggExec(PL.HashAgg, [E.GroupingExpr("k", E.Column("k"))],
                                    [E.AggExpr(nm, E.PARTIAL, PL.create_agg(fn, [E.Column("v")], sch, rt)) for nm, fn, rt in specs], True, PL.MemoryExec(sch))
run("M1 f64 SUM+COUNT (wide tile)", mkp(s1f, [("s", E.AGG_SUM, T.float64), ("c", E.AGG_COUNT, T.int64)]).plan_bytes(), [k, vf], 16.0, native.default_conf(agg_initial_groups=1 << 20))
run("M1 f64 AVG (wide tile)", mkp(s1f, [("a", E.AGG_AVG, T.float64)]).plan_bytes(), [k, vf], 16.0, native.default_conf(agg_initial_groups=1 << 20))
run("M1 f64 SUM+COUNT generic VM", mkp(s1f, [("s", E.AGG_SUM, T.float64), ("c", E.AGG_COUNT, T.int64)]).plan_bytes(), [k, vf], 16.0, native.default_conf(agg_initial_groups=1 << 20, force_generic_kernels=1), reps=1)
del vf
run("M1 int64 MIN+MAX (wide tile)", mkp(s1, [("mn", E.AGG_MIN, T.int64), ("mx", E.AGG_MAX, T.int64)]).plan_bytes(), [k, v], 16.0, native.default_conf(agg_initial_groups=1 << 20))
vd = torch.stack([v, v >> 63], dim=1).contiguous()          # decimal128(17,2): little-endian {lo, hi} pairs, sign-extended
s1d = T.Schema([T.Field("k", T.int64, False), T.Field("v", T.decimal128(17, 2), False)])
run("M1 decimal128(17,2) SUM+COUNT (wide tile)", mkp(s1d, [("s", E.AGG_SUM, T.decimal128(27, 2)), ("c", E.AGG_COUNT, T.int64)]).plan_bytes(), [k, vd], 24.0, native.default_conf(agg_initial_groups=1 << 20))
run("M1 decimal128(17,2) SUM+COUNT generic VM", mkp(s1d, [("s", E.AGG_SUM, T.decimal128(27, 2)), ("c", E.AGG_COUNT, T.int64)]).plan_bytes(), [k, vd], 24.0, native.default_conf(agg_initial_groups=1 << 20, force_generic_kernels=1), reps=1)
del vd
run("M1 generic VM kernel", m1.plan_bytes(), [k, v], 16.0, native.default_conf(agg_initial_groups=1 << 20, force_generic_kernels=1), reps=1)
del k
# M2: q1-shaped: f BETWEEN lo AND hi (s = 0.2), keys (k1 ~ U[0,2^17), k2 ~ U[0,8)), SUM(v)   (32 B/row)
f = torch.randint(0, 1000, (rows,), dtype=torch.int64, device=dev, generator=g)
k1 = torch.randint(0, 1 << 17, (rows,), dtype=torch.int64, device=dev, generator=g)
k2 = torch.randint(0, 8, (rows,), dtype=torch.int64, device=dev, generator=g)
s2 = T.Schema([T.Field(n, T.int64, False) for n in ("f", "k1", "k2", "v")])
preds = [E.BinaryExpr(E.Column("f"), "GtEq", E.Literal(200, T.int64)), E.BinaryExpr(E.Column("f"), "LtEq", E.Literal(399, T.int64))]
m2 = PL.AggExec(PL.HashAgg, [E.GroupingExpr("k1", E.Column("k1")), E.GroupingExpr("k2", E.Column("k2"))],
                [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], s2, T.int64))], True, PL.FilterExec(preds, PL.MemoryExec(s2)))
run("M2 q1-shaped fused filter->agg (2 keys)", m2.plan_bytes(), [f, k1, k2, v], 32.0, native.default_conf(agg_initial_groups=1 << 20))
run("M2 q1-shaped", m2.plan_bytes(), [f, k1, k2, v], 32.0, native.default_conf(agg_initial_groups=1 << 20), reps=1, steady=True)
# M3: ShuffleWriterExec 200-way hash partition + batch_serde encode (BASELINE configs[3] map side): 32 B/row read + 32 B/row of byte planes written
m3 = PL.ShuffleWriterExec(PL.MemoryExec(s2), ("hash", [E.Column("k1")], 200), "", "")
run("M3 shuffle write 200-way (4 int64 columns, hash on k1)", m3.plan_bytes(), [f, k1, k2, v], 64.0, native.default_conf(shuffle_output_on_device=1), reps=2)
m3b = PL.ShuffleWriterExec(PL.MemoryExec(s2), ("hash", [E.Column("k1"), E.Column("k2")], 2000), "", "")
run("M3 shuffle write 2000-way (hash on k1,k2)", m3b.plan_bytes(), [f, k1, k2, v], 64.0, native.default_conf(shuffle_output_on_device=1), reps=2)


# M4: HashJoinExec store_sales JOIN date_dim (BASELINE configs[3] probe side): the map side is its own op, the probed side streams.
def run_join(name, n_build_keep, alg_bytes_per_row):
    if ONLY and not any(t in name for t in ONLY.split(",")): return
    ND = 73049                                                     # rows of TPC-DS date_dim
    d_sk = torch.arange(ND, dtype=torch.int64, device=dev)
    d_year = 1900 + d_sk // 366
    d_moy = (d_sk // 30) % 12 + 1
    keep = d_sk < n_build_keep                                      # the dimension filter (e.g. one year) already applied to the map side
    bcols = [d_sk[keep].contiguous(), d_year[keep].contiguous(), d_moy[keep].contiguous()]
    sk = torch.randint(0, ND, (rows,), dtype=torch.int64, device=dev, generator=g)
    sd = T.Schema([T.Field(n, T.int64, False) for n in ("d_date_sk", "d_year", "d_moy")])
    ss = T.Schema([T.Field(n, T.int64, False) for n in ("ss_sold_date_sk", "ss_item_sk", "ss_quantity", "ss_net_paid")])
    build = PL.BroadcastJoinBuildHashMapExec(PL.MemoryExec(sd), [E.Column("d_date_sk")])
    join = PL.BroadcastJoinExec(PL.build_join_schema(ss, sd, PL.JOIN_INNER), PL.MemoryExec(ss), build, [(E.Column("ss_sold_date_sk"), E.Column("d_date_sk"))], PL.JOIN_INNER, PL.RIGHT_SIDE, True, "m")
    nb = int(bcols[0].numel())
    best = None
    for _ in range(REPS or 3):
        with native.NativeOp(build.plan_bytes(), native.default_conf(), 0) as bop:
            bop.push_device(native.DeviceBatch([(c.data_ptr(), 0, nb) for c in bcols], nb, 0, keepalive=tuple(bcols)))
            bop.finish()
            with native.NativeOp(join.plan_bytes(), native.default_conf(), 0) as op:
                op.attach_build(bop)
                pc = [sk, k1, k2, v]
                op.push_device(native.DeviceBatch([(c.data_ptr(), 0, rows) for c in pc], rows, 0, keepalive=tuple(pc)))
                op.finish()
                n_out = 0
                while True:
                    o = op.pull_device()
                    if o is None: break
                    n_out += o.array.length; native.release_device_array(o)
                m = op.metrics()
        if best is None or m["hot_kernel_ns"] < best[0]: best = (m["hot_kernel_ns"], m, n_out)
    t, m, n_out = best
    gbs = alg_bytes_per_row * rows / t
    print(json.dumps({"shape": name, "rows": rows, "build_rows": nb, "out_rows": n_out, "probe_ms": t / 1e6, "rows_per_s": rows / (t * 1e-9), "alg_GBps": gbs, "frac_of_measured_hbm": gbs / peak,
                      "launches": m["gpu_kernel_launches"]}), flush=True)

run_join("M4 hash join store_sales x date_dim, every row matches (32 B read + 56 B written per probe row)", 73049, 88.0)
run_join("M4 hash join store_sales x date_dim filtered to one year (0.5 % match; 32 B read per probe row)", 366, 32.0 + 0.005 * 56)


# M5: SortExec ORDER BY one int64 key (full 64-bit range: all eight radix passes) over (key, payload), and a 20-bit key (three passes)
def run_sort(name, key, n_sort):
    if ONLY and not any(t in name for t in ONLY.split(",")): return
    ssch = T.Schema([T.Field("k", T.int64, False), T.Field("p", T.int64, False)])
    plan = PL.SortExec(PL.MemoryExec(ssch), [(E.Column("k"), False, True)])
    kk, pp = key[:n_sort].contiguous(), v[:n_sort].contiguous()
    best = None
    for _ in range(REPS or 2):
        with native.NativeOp(plan.plan_bytes(), native.default_conf(), 0) as op:
            op.push_device(native.DeviceBatch([(kk.data_ptr(), 0, n_sort), (pp.data_ptr(), 0, n_sort)], n_sort, 0, keepalive=(kk, pp)))
            op.finish()
            o = op.pull_device()
            srt = torch.as_tensor(__import__("bench").CudaView(o.array.children[0].contents.buffers[1], n_sort * 8, o), device="cuda").view(torch.int64)
            ok = bool((srt[1:] >= srt[:-1]).all())
            native.release_device_array(o)
            m = op.metrics()
        if best is None or m["hot_kernel_ns"] < best[0]: best = (m["hot_kernel_ns"], m, ok)
    t, m, ok = best
    print(json.dumps({"shape": name, "rows": n_sort, "sorted": ok, "sort_ms": t / 1e6, "rows_per_s": n_sort / (t * 1e-9), "launches": m["gpu_kernel_launches"]}), flush=True)

run_sort("M5 sort int64 key, full range (8 radix passes) + 8-byte payload", (k1 * 2654435761 * 40503 + f * 2**40) ^ (v << 20), min(rows, 1 << 26))
run_sort("M5 sort int64 key in [0, 2^17) (3 radix passes) + 8-byte payload", k1, min(rows, 1 << 26))


# M6: ParquetScanExec (BASELINE configs[2] shape): store_sales-like synthetic columns written by pyarrow, scanned (decode on the GPU) with a
# pushed-down predicate on the sorted date key; host work (file read, Thrift, Snappy) is inside the measured time
def run_parquet(name, compression, n_pq):
    if ONLY and not any(t in name for t in ONLY.split(",")): return
    import time, tempfile, numpy as np, pyarrow as pa, pyarrow.parquet as pq
    rng = np.random.default_rng(7)
    tbl = pa.table({"ss_sold_date_sk": pa.array(np.sort(rng.integers(2450816, 2452642, n_pq)).astype(np.int32), pa.int32()),
                    "ss_item_sk": pa.array(rng.integers(1, 204000, n_pq).astype(np.int32), pa.int32()),
                    "ss_quantity": pa.array(rng.integers(1, 100, n_pq).astype(np.int32), pa.int32()),
                    "ss_net_paid": pa.array(rng.integers(0, 2000000, n_pq, dtype=np.int64))})
    tbl = tbl.cast(pa.schema([pa.field(f.name, f.type, False) for f in tbl.schema]))
    path = os.path.join(tempfile.gettempdir(), f"b200q_store_sales_{compression}.parquet")
    pq.write_table(tbl, path, compression=compression, row_group_size=1 << 20)
    fsz = os.path.getsize(path)
    sch = T.from_arrow_schema(tbl.schema)
    pred = E.BinaryExpr(E.Column("ss_sold_date_sk"), "GtEq", E.Literal(2452000, T.int32))
    for label, preds in (("full scan", []), ("pushdown ss_sold_date_sk >= 2452000", [pred])):
        scan = PL.ParquetScanExec(sch, [(path, fsz, None)], pruning_predicates=preds)
        plan = PL.FilterExec(preds, scan) if preds else scan
        best = None
        for _ in range(3):
            pb = plan.plan_bytes()
            t0 = time.perf_counter()
            with native.NativeOp(pb, native.default_conf(), 0) as op:
                t1 = time.perf_counter()
                op.finish()
                t2 = time.perf_counter()
                n_out = 0
                while True:
                    o = op.pull_device()
                    if o is None: break
                    n_out += o.array.length; native.release_device_array(o)
                m = op.metrics()
                t3 = time.perf_counter()
            dt = time.perf_counter() - t0
            m["phases_ms"] = {"create": (t1 - t0) * 1e3, "finish (the scan)": (t2 - t1) * 1e3, "pull_device + release": (t3 - t2) * 1e3, "destroy": (t0 + dt - t3) * 1e3}
            if best is None or dt < best[0]: best = (dt, n_out, m)
        dt, n_out, m = best
        print(json.dumps({"shape": f"{name} [{label}]", "rows": n_pq, "file_bytes": fsz, "out_rows": n_out, "wall_ms": dt * 1e3, "rows_per_s": n_pq / dt, "file_GBps": fsz / dt / 1e9,
                          "gpu_ms": m["elapsed_compute_ns"] / 1e6, "row_groups_decoded": m["input_batches"], "row_groups_pruned": m["fast_path_launches"], "launches": m["gpu_kernel_launches"], "phases_ms": m["phases_ms"]}), flush=True)
    t0 = time.perf_counter(); pq.read_table(path); print(json.dumps({"shape": f"{name} [pyarrow.parquet.read_table on the host, all cores]", "wall_ms": (time.perf_counter() - t0) * 1e3}), flush=True)

run_parquet("M6 parquet scan store_sales-like 4 columns, snappy + dictionary", "snappy", 1 << 24)
run_parquet("M6 parquet scan store_sales-like 4 columns, uncompressed", "none", 1 << 24)
