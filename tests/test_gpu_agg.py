"""HashAggregateExec on the GPU vs the oracle (multiset parity; bit-exact ints/decimals, 1e-6 fp64)."""
import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, plans as PL, types as T, native
from oracle import blaze_oracle as O
from helpers import *

pytestmark = pytest.mark.gpu


def agg_exprs(mode, specs, ins):
    return [E.AggExpr(name, mode, PL.create_agg(fn, children, ins, rt)) for name, fn, children, rt in specs]


def run_partial_final(rb, group_cols, specs, batch_rows=10000, conf=None, float_cols=(), columnar=False):
    """Partial -> Final through the reference's Binary agg-buffer column (or the columnar state),
    both stages fused in one op; compared with the oracle running the same two AggExecs."""
    batches = split_batches(rb, batch_rows)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    groupings = [E.GroupingExpr(c, E.Column(c)) for c in group_cols]
    partial = PL.AggExec(PL.HashAgg, groupings, agg_exprs(PL.Partial, specs, ins), False, leaf, columnar_state=columnar)
    specs_final = [(name, fn, [E.placeholder(ch[0].data_type(ins))] if ch else [], rt) for name, fn, ch, rt in specs]
    final = PL.AggExec(PL.HashAgg, groupings, agg_exprs(PL.Final, specs_final, partial.schema()), False, partial)
    if columnar:
        conf = conf or native.default_conf()
        conf.partial_state_columnar = 1
    got = PL.collect(final, conf)
    ob = oracle_batches(batches)
    o_partial = O.AggExec(E.HASH_AGG, groupings, agg_exprs(E.PARTIAL, specs, ins), False, ins)
    mid = o_partial.execute(ob)
    o_final = O.AggExec(E.HASH_AGG, groupings, agg_exprs(E.FINAL, specs_final, o_partial.schema), False, o_partial.schema)
    exp = o_final.execute(mid)
    assert_multiset_equal(got, exp, float_cols)
    return got, final


def test_reference_kat_test_agg():
    """the reference's own golden table (agg_exec.rs:493-681), in-scope aggregates"""
    cols = {"a": [2, 9, 3, 1, 0, 4, 6], "b": [1, 0, 0, 3, 5, 6, 3], "c": [7, 8, 7, 8, 9, 2, 5], "d": [-7, 86, 71, 83, 90, -2, 5],
            "e": [-7, 86, 71, 83, 90, -2, 5], "f": [0, 1, 2, 3, 4, 5, 6], "g": [6, 3, 6, 3, 1, 5, 4], "h": [6, 3, 6, 3, 1, 5, 4]}
    rb = rb_from_cols(list(cols), [pa.array(v, pa.int32()) for v in cols.values()])
    specs = [("agg_expr_sum", E.AGG_SUM, [E.Column("a")], T.int64), ("agg_expr_avg", E.AGG_AVG, [E.Column("b")], T.float64),
             ("agg_expr_max", E.AGG_MAX, [E.Column("d")], T.int32), ("agg_expr_min", E.AGG_MIN, [E.Column("e")], T.int32),
             ("agg_expr_count", E.AGG_COUNT, [E.Column("f")], T.int64)]
    got, _ = run_partial_final(rb, ["c"], specs)
    t = pa.Table.from_batches(got).sort_by("c").to_pydict()
    assert t == {"c": [2, 5, 7, 8, 9], "agg_expr_sum": [4, 6, 5, 10, 0], "agg_expr_avg": [6.0, 3.0, 0.5, 1.5, 5.0],
                 "agg_expr_max": [-2, 5, 71, 86, 90], "agg_expr_min": [-2, 5, -7, 83, 90], "agg_expr_count": [1, 1, 2, 2, 1]}


def m1_batch(n, card, seed=44, null_frac=0.0, key_null_frac=0.0):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, card, n, dtype=np.int64)
    v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    return rb_from_cols(["k", "v"], [with_nulls(rng, k, key_null_frac), with_nulls(rng, v, null_frac)])


@pytest.mark.parametrize("columnar", [False, True])
@pytest.mark.parametrize("n,card,nf,knf", [(1, 1, 0, 0), (1000, 7, 0.3, 0.1), (200_000, 50_000, 0.1, 0.0), (300_000, 3, 0.0, 0.05)])
def test_m1_sum_count(n, card, nf, knf, columnar):
    rb = m1_batch(n, card, 44, nf, knf)
    specs = [("s", E.AGG_SUM, [E.Column("v")], T.int64), ("c", E.AGG_COUNT, [E.Column("v")], T.int64),
             ("mn", E.AGG_MIN, [E.Column("v")], T.int64), ("mx", E.AGG_MAX, [E.Column("v")], T.int64),
             ("av", E.AGG_AVG, [E.Column("v")], T.float64)]
    run_partial_final(rb, ["k"], specs, float_cols=(5,), columnar=columnar)


def test_all_null_group_and_wrapping_sum():
    k = pa.array([1, 1, 2, 2, 3], pa.int64())
    v = pa.array([2**62, 2**62, None, None, -5], pa.int64())
    rb = rb_from_cols(["k", "v"], [k, v])
    specs = [("s", E.AGG_SUM, [E.Column("v")], T.int64), ("c", E.AGG_COUNT, [E.Column("v")], T.int64)]
    got, _ = run_partial_final(rb, ["k"], specs)
    t = pa.Table.from_batches(got).sort_by("k").to_pydict()
    assert t["s"] == [-2**63, None, -5] and t["c"] == [2, 0, 1]      # wrapping add (Cargo.toml:43-45), NULL for an all-NULL group


def test_table_growth_many_groups():
    n = 600_000
    rb = m1_batch(n, 2**40, 45)                                       # ~all keys distinct: forces rehash + replay of deferred rows
    specs = [("s", E.AGG_SUM, [E.Column("v")], T.int64), ("c", E.AGG_COUNT, [E.Column("v")], T.int64)]
    conf = native.default_conf(agg_initial_groups=1024, staging_rows=0)
    got, plan = run_partial_final(rb, ["k"], specs, batch_rows=n, conf=conf)
    assert plan.last_metrics["table_grow_count"] >= 0


@pytest.mark.parametrize("n,launch_rows", [(1_500_000, 1 << 16), (700_000, 200_000)])
def test_table_growth_while_the_next_launch_is_in_flight(n, launch_rows):
    # a batch is cut into launches that are enqueued back to back; the counters of launch i are read while launch i + 1 runs, so when the
    # table hits its load limit TWO launches hold deferred rows: both are replayed (stages.cu update_rows / settle)
    rb = m1_batch(n, 2**40, 47)
    specs = [("s", E.AGG_SUM, [E.Column("v")], T.int64), ("c", E.AGG_COUNT, [E.Column("v")], T.int64)]
    conf = native.default_conf(agg_initial_groups=1024, staging_rows=0, max_launch_rows=launch_rows)
    got, plan = run_partial_final(rb, ["k"], specs, batch_rows=n, conf=conf)
    assert plan.last_metrics["table_grow_count"] >= 1


def test_two_keys_filter_fused_q1_shape():
    n = 250_000
    rng = np.random.default_rng(46)
    f = rng.integers(0, 1000, n, dtype=np.int64)
    k1 = rng.integers(0, 2**10, n, dtype=np.int64)
    k2 = rng.integers(0, 8, n).astype(np.int32)
    v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    rb = rb_from_cols(["f", "k1", "k2", "v"], [pa.array(f), with_nulls(rng, k1, 0.01), pa.array(k2), with_nulls(rng, v, 0.05)])
    batches = split_batches(rb, 10000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    preds = [E.BinaryExpr(E.Column("f"), "GtEq", E.Literal(200, T.int64)), E.BinaryExpr(E.Column("f"), "LtEq", E.Literal(399, T.int64))]
    filt = PL.FilterExec(preds, leaf)
    groupings = [E.GroupingExpr("k1", E.Column("k1")), E.GroupingExpr("k2", E.Column("k2"))]
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], ins, T.int64))]
    plan = PL.AggExec(PL.HashAgg, groupings, aggs, True, filt)
    got = PL.collect(plan)
    ob = oracle_batches(batches)
    exp = O.AggExec(E.HASH_AGG, groupings, aggs, False, ins).execute(O.FilterExec(preds, ins).execute(ob))
    assert_multiset_equal(got, exp)                                   # includes the frozen Binary column, byte for byte


def test_no_grouping_and_empty_input():
    rb = m1_batch(5000, 10, 47, 0.2)
    leaf = PL.MemoryExec.from_arrow(split_batches(rb, 1000), rb.schema)
    ins = leaf.schema()
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], ins, T.int64)),
            E.AggExpr("c", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Column("v")], ins, T.int64))]
    plan = PL.AggExec(PL.HashAgg, [], aggs, False, leaf)
    got = PL.collect(plan)
    exp = O.AggExec(E.HASH_AGG, [], aggs, False, ins).execute(oracle_batches(leaf.batches))
    assert_multiset_equal(got, exp)
    empty = PL.MemoryExec(ins, [])
    got = PL.collect(PL.AggExec(PL.HashAgg, [], aggs, False, empty))
    exp = O.AggExec(E.HASH_AGG, [], aggs, False, ins).execute([])
    assert_multiset_equal(got, exp)
    assert PL.collect(PL.AggExec(PL.HashAgg, [E.GroupingExpr("k", E.Column("k"))], aggs, False, empty)) == []


def test_no_grouping_over_non_nullable_arguments_yields_null_without_rows():
    """execute_agg_no_grouping (agg_exec.rs:280-323) always emits one row; with no row reaching the accumulators
    their valids stay false: SUM/MIN/MAX/AVG = NULL, COUNT = 0 — also when the argument column is declared non-null."""
    n = 3000
    rng = np.random.default_rng(77)
    schema = pa.schema([pa.field("f", pa.int64(), nullable=False), pa.field("v", pa.int64(), nullable=False)])
    rb = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 100, n, dtype=np.int64)), pa.array(rng.integers(-50, 50, n, dtype=np.int64))], schema=schema)
    specs = [("s", E.AGG_SUM, T.int64), ("mn", E.AGG_MIN, T.int64), ("mx", E.AGG_MAX, T.int64), ("a", E.AGG_AVG, T.float64), ("c", E.AGG_COUNT, T.int64)]
    for batches, preds in (([], []), (split_batches(rb, 1000), [E.BinaryExpr(E.Column("f"), "Lt", E.Literal(-1, T.int64))]),
                           (split_batches(rb, 1000), [E.BinaryExpr(E.Column("f"), "Lt", E.Literal(50, T.int64))])):
        leaf = PL.MemoryExec.from_arrow(batches, schema)
        ins = leaf.schema()
        mk = lambda mode, ch: [E.AggExpr(nm, mode, PL.create_agg(fn, ch(fn, rt), ins, rt)) for nm, fn, rt in specs]
        child = PL.FilterExec(preds, leaf) if preds else leaf
        partial = PL.AggExec(PL.HashAgg, [], mk(E.PARTIAL, lambda fn, rt: [E.Column("v")]), False, child)
        final = PL.AggExec(PL.HashAgg, [], mk(E.FINAL, lambda fn, rt: [E.placeholder(rt)]), False, partial)
        got_p = PL.collect(partial)
        got = PL.collect(final)
        ob = oracle_batches(batches)
        op = O.AggExec(E.HASH_AGG, [], mk(E.PARTIAL, lambda fn, rt: [E.Column("v")]), False, ins)
        of = O.AggExec(E.HASH_AGG, [], mk(E.FINAL, lambda fn, rt: [E.placeholder(rt)]), False, op.schema)
        exp_p = op.execute(O.FilterExec(preds, ins).execute(ob) if preds else ob)
        assert_multiset_equal(got_p, exp_p)                           # the frozen state bytes carry valid = 0
        assert_multiset_equal(got, of.execute(exp_p), float_cols=(3,))
        if not batches or preds[0].right.value == -1:
            row = got[0].to_pylist()[0]
            assert [row[k] for k in ("s", "mn", "mx", "a", "c")] == [None, None, None, None, 0]


def test_table_over_the_hbm_budget_is_unsupported_not_a_cuda_failure():
    """A12 (agg_table.rs:108-120,540-588 is what the reference does under memory pressure): the GPU table cannot
    spill, so outgrowing b200q_conf.agg_max_table_bytes is B200Q_ERR_UNSUPPORTED (host falls back);
    the handle stays usable for teardown and the same plan runs when the budget allows it."""
    n = 600_000
    rng = np.random.default_rng(5)
    k = rng.integers(0, 2**40, n, dtype=np.int64)                    # ~600k sparse groups: hashed table, must grow past 2^20 slots
    rb = rb_from_cols(["k", "v"], [pa.array(k), pa.array(rng.integers(-9, 9, n, dtype=np.int64))])
    batches = split_batches(rb, 100_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], ins, T.int64))]
    plan = PL.AggExec(PL.HashAgg, [E.GroupingExpr("k", E.Column("k"))], aggs, False, leaf)
    with pytest.raises(native.NativeError) as ei:
        PL.collect(plan, native.default_conf(staging_rows=0, agg_initial_groups=1 << 10, agg_max_table_bytes=40 << 20))   # 2^20 slots x 24 B fit, 2^21 do not
    assert ei.value.code == native.ERR_UNSUPPORTED and "HBM budget" in str(ei.value)
    got = PL.collect(plan, native.default_conf(staging_rows=0, agg_initial_groups=1 << 10, agg_max_table_bytes=1 << 30))
    exp = O.AggExec(E.HASH_AGG, [E.GroupingExpr("k", E.Column("k"))], aggs, False, ins).execute(oracle_batches(batches))
    assert_multiset_equal(got, exp)
    assert plan.last_metrics["table_grow_count"] >= 1


def test_f64_and_decimal_sums():
    import decimal
    n = 120_000
    rng = np.random.default_rng(48)
    k = rng.integers(0, 5000, n, dtype=np.int64)
    x = rng.normal(0, 1e6, n)
    raw = rng.integers(-10**15, 10**15, n)
    dec = pa.array([decimal.Decimal(int(r)).scaleb(-2) for r in raw], type=pa.decimal128(17, 2))
    rb = rb_from_cols(["k", "x", "d"], [pa.array(k), with_nulls(rng, x, 0.1), dec])
    specs = [("sx", E.AGG_SUM, [E.Column("x")], T.float64), ("ax", E.AGG_AVG, [E.Column("x")], T.float64),
             ("mnx", E.AGG_MIN, [E.Column("x")], T.float64), ("mxx", E.AGG_MAX, [E.Column("x")], T.float64),
             ("sd", E.AGG_SUM, [E.Column("d")], T.decimal128(27, 2)), ("ad", E.AGG_AVG, [E.Column("d")], T.decimal128(21, 6)),
             ("mnd", E.AGG_MIN, [E.Column("d")], T.decimal128(17, 2)), ("mxd", E.AGG_MAX, [E.Column("d")], T.decimal128(17, 2)),
             ("c", E.AGG_COUNT, [E.Column("x")], T.int64)]
    run_partial_final(rb, ["k"], specs, float_cols=(1, 2))


MODES = {"fast+dense": {}, "fast-hash": {"agg_dense_keys": 0}, "generic": {"force_generic_kernels": 1}}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("n,card,nf,knf,lo", [(1, 1, 0, 0, 0), (4099, 17, 0.3, 0.2, -5), (400_000, 100_000, 0.0, 0.0, 10**12), (300_000, 2**45, 0.1, 0.01, 0)])
def test_fast_paths_sum_count(mode, n, card, nf, knf, lo):
    """the specialised kernels (lane-paired REDs, dense direct indexing) vs the generic VM kernel vs the oracle"""
    rng = np.random.default_rng(50)
    k = rng.integers(lo, lo + card, n, dtype=np.int64)
    v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    rb = rb_from_cols(["k", "v"], [with_nulls(rng, k, knf), with_nulls(rng, v, nf)])
    specs = [("s", E.AGG_SUM, [E.Column("v")], T.int64), ("c", E.AGG_COUNT, [E.Column("v")], T.int64)]
    conf = native.default_conf(staging_rows=0, **MODES[mode])
    got, plan = run_partial_final(rb, ["k"], specs, batch_rows=150_000, conf=conf)
    m = plan.last_metrics
    if mode == "generic":
        assert m["fast_path_launches"] == 0
    else:
        assert m["fast_path_launches"] > 0


@pytest.mark.parametrize("mode", list(MODES))
def test_fast_paths_filters_two_keys_small_ints(mode):
    n = 200_000
    rng = np.random.default_rng(51)
    d = pa.array(rng.integers(0, 3000, n, dtype=np.int32), pa.int32()).cast(pa.date32())
    k1 = rng.integers(0, 300, n).astype(np.int16)
    k2 = rng.integers(-3, 3, n).astype(np.int8)
    v = rng.integers(-1000, 1000, n).astype(np.int32)
    rb = rb_from_cols(["d", "k1", "k2", "v"], [d, with_nulls(rng, k1, 0.02, pa.int16()), pa.array(k2), with_nulls(rng, v, 0.1, pa.int32())])
    batches = split_batches(rb, 10000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    preds = [E.BinaryExpr(E.Column("d"), "GtEq", E.Literal(1000, T.date32)), E.BinaryExpr(E.Literal(2000, T.date32), "Gt", E.Column("d")),
             E.BinaryExpr(E.Column("k2"), "NotEq", E.Literal(0, T.int8))]
    groupings = [E.GroupingExpr("k1", E.Column("k1")), E.GroupingExpr("k2", E.Column("k2"))]
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], ins, T.int64)),
            E.AggExpr("n", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Literal(1, T.int64)], ins, T.int64))]
    plan = PL.AggExec(PL.HashAgg, groupings, aggs, True, PL.FilterExec(preds, leaf))
    got = PL.collect(plan, native.default_conf(**MODES[mode]))
    exp = O.AggExec(E.HASH_AGG, groupings, aggs, False, ins).execute(O.FilterExec(preds, ins).execute(oracle_batches(batches)))
    assert_multiset_equal(got, exp)
    assert (plan.last_metrics["fast_path_launches"] > 0) == (mode != "generic")


def test_large_batch_against_c_port():
    """64M rows (BASELINE M1 shape, one device-sized host batch): compared with the C restatement of the reference
    algorithm (oracle/cpu_ref.c) + a checksum-of-checksums property independent of any oracle"""
    from oracle import cpu_ref
    n = 1 << 26
    rng = np.random.default_rng(52)
    k = rng.integers(0, 1 << 20, n, dtype=np.int64)
    v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    rb = rb_from_cols(["k", "v"], [pa.array(k), pa.array(v)])
    leaf = PL.MemoryExec.from_arrow([rb], rb.schema)
    ins = leaf.schema()
    g = [E.GroupingExpr("k", E.Column("k"))]
    mk = lambda mode, ch: [E.AggExpr("s", mode, PL.create_agg(E.AGG_SUM, ch, ins, T.int64)), E.AggExpr("c", mode, PL.create_agg(E.AGG_COUNT, ch, ins, T.int64))]
    final = PL.AggExec(PL.HashAgg, g, mk(E.FINAL, [E.placeholder(T.int64)]), False, PL.AggExec(PL.HashAgg, g, mk(E.PARTIAL, [E.Column("v")]), False, leaf))
    t = pa.Table.from_batches(PL.collect(final, native.default_conf(staging_rows=0, batch_size=1 << 20)))
    gk, gs, gc = t["k"].to_numpy(), t["s"].to_numpy(), t["c"].to_numpy()
    assert int(gc.sum()) == n and int(gs.sum()) == int(v.sum()) and len(np.unique(gk)) == len(gk)
    ref = cpu_ref.hashagg_sum_count(k, v, nthreads=8, max_groups=1 << 21)
    order_g, order_r = np.argsort(gk), np.argsort(ref["k"])
    assert np.array_equal(gk[order_g], ref["k"][order_r]) and np.array_equal(gs[order_g], ref["sum"][order_r]) and np.array_equal(gc[order_g], ref["count"][order_r])


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("shape", ["lean-global", "lean-smem", "typed", "sum-only-notnull"])
def test_two_key_composite_dense_index(mode, shape):
    """two small-range integer keys are mapped onto ONE dense index (k0-b0)*r1 + (k1-b1); later batches bring keys
    outside the range decided on the first batch (-> hashed slots) — every form must agree with the oracle"""
    n = 300_000
    rng = np.random.default_rng(77)
    r0 = 40 if shape == "lean-smem" else 6000
    k0 = rng.integers(100, 100 + r0, n, dtype=np.int64); k0[200_000:] += rng.integers(0, 3 * r0, n - 200_000)     # range grows after batch 1
    k1 = rng.integers(-3, 4, n, dtype=np.int64); k1[250_000:] = rng.integers(-40, 40, n - 250_000)
    v = rng.integers(-10**9, 10**9, n, dtype=np.int64)
    if shape == "typed":
        cols = [with_nulls(rng, k0.astype(np.int32), 0.01, pa.int32()), with_nulls(rng, k1.astype(np.int16), 0.02, pa.int16()), with_nulls(rng, v, 0.2)]
        rb = rb_from_cols(["k0", "k1", "v"], cols)
    else:
        notnull = shape == "sum-only-notnull"
        schema = pa.schema([pa.field("k0", pa.int64(), nullable=not notnull), pa.field("k1", pa.int64(), nullable=not notnull), pa.field("v", pa.int64(), nullable=not notnull)])
        rb = pa.RecordBatch.from_arrays([pa.array(k0), pa.array(k1), pa.array(v)], schema=schema)
    specs = [("s", E.AGG_SUM, [E.Column("v")], T.int64)] if shape == "sum-only-notnull" else \
            [("s", E.AGG_SUM, [E.Column("v")], T.int64), ("c", E.AGG_COUNT, [E.Column("v")], T.int64)]
    conf = native.default_conf(staging_rows=0, **MODES[mode])
    got, plan = run_partial_final(rb, ["k0", "k1"], specs, batch_rows=100_000, conf=conf)
    assert (plan.last_metrics["fast_path_launches"] > 0) == (mode != "generic")


@pytest.mark.parametrize("mode", list(MODES))
def test_one_key_fused_filter_lean_dense_row_form(mode):
    """non-null int64 inputs + fused conjuncts + {SUM, COUNT(*)}: 2-word dense entries updated by lane pairs (one row per lane)"""
    n = 250_001
    rng = np.random.default_rng(78)
    schema = pa.schema([pa.field(c, pa.int64(), nullable=False) for c in ("k", "f", "v")])
    k = rng.integers(-500, 70_000, n, dtype=np.int64); k[1::2] = k[::2][: n // 2]          # neighbouring rows often share a group
    rb = pa.RecordBatch.from_arrays([pa.array(k), pa.array(rng.integers(0, 100, n, dtype=np.int64)), pa.array(rng.integers(-2**40, 2**40, n, dtype=np.int64))], schema=schema)
    batches = split_batches(rb, 100_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    preds = [E.BinaryExpr(E.Column("f"), "GtEq", E.Literal(20, T.int64)), E.BinaryExpr(E.Column("f"), "Lt", E.Literal(55, T.int64))]
    groupings = [E.GroupingExpr("k", E.Column("k"))]
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], ins, T.int64)),
            E.AggExpr("n", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Literal(1, T.int64)], ins, T.int64))]
    plan = PL.AggExec(PL.HashAgg, groupings, aggs, True, PL.FilterExec(preds, leaf))
    got = PL.collect(plan, native.default_conf(staging_rows=0, **MODES[mode]))
    exp = O.AggExec(E.HASH_AGG, groupings, aggs, False, ins).execute(O.FilterExec(preds, ins).execute(oracle_batches(batches)))
    assert_multiset_equal(got, exp)
    assert (plan.last_metrics["fast_path_launches"] > 0) == (mode != "generic")
