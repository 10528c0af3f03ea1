"""The C restatement (oracle/cpu_ref.c, the timed CPU baseline) must agree with the numpy oracle."""
import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, types as T
from oracle import blaze_oracle as O
from oracle import cpu_ref


@pytest.mark.parametrize("nthreads", [1, 3])
@pytest.mark.parametrize("n,card,nf,knf", [(0, 1, 0, 0), (1, 1, 0, 0), (25_000, 1000, 0.2, 0.05), (60_000, 2**40, 0.0, 0.0)])
def test_hashagg_port_matches_numpy_oracle(n, card, nf, knf, nthreads):
    rng = np.random.default_rng(5)
    k = rng.integers(0, card, n, dtype=np.int64); v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    kv = rng.random(n) >= knf; vv = rng.random(n) >= nf
    got = cpu_ref.hashagg_sum_count(k, v, kv, vv, nthreads)
    ins = T.Schema([T.Field("k", T.int64, True), T.Field("v", T.int64, True)])
    b = O.Batch(ins, [O.Col(T.int64, k, kv), O.Col(T.int64, v, vv)], n)
    g = [E.GroupingExpr("k", E.Column("k"))]
    aggs = [E.AggExpr("s", E.PARTIAL, E.AggFunctionExpr(E.AGG_SUM, [E.Column("v")], T.int64)),
            E.AggExpr("c", E.PARTIAL, E.AggFunctionExpr(E.AGG_COUNT, [E.Column("v")], T.int64))]
    batches = [b.take(np.arange(i, min(i + 10000, n))) for i in range(0, n, 10000)]
    part = O.AggExec(E.HASH_AGG, g, aggs, False, ins)
    fin = O.AggExec(E.HASH_AGG, g, [E.AggExpr("s", E.FINAL, E.AggFunctionExpr(E.AGG_SUM, [E.placeholder(T.int64)], T.int64)),
                                    E.AggExpr("c", E.FINAL, E.AggFunctionExpr(E.AGG_COUNT, [E.placeholder(T.int64)], T.int64))], False, part.schema)
    exp = O.rows_multiset(fin.execute(part.execute(batches)))
    ms = {}
    for i in range(len(got["k"])):
        key = (int(got["k"][i]) if got["k_valid"][i] else None, int(got["sum"][i]) if got["sum_valid"][i] else None, int(got["count"][i]))
        ms[key] = ms.get(key, 0) + 1
    assert ms == exp


@pytest.mark.parametrize("nthreads", [1, 4])
def test_filter_project_port_matches_numpy_oracle(nthreads):
    rng = np.random.default_rng(6)
    n = 45_001
    a = rng.integers(0, 1000, n, dtype=np.int64); b = rng.integers(-2**31, 2**31, n, dtype=np.int64)
    av = rng.random(n) >= 0.1; bv = rng.random(n) >= 0.1
    oa, oc, ocv = cpu_ref.filter_project(a, b, 500, av, bv, nthreads)
    ins = T.Schema([T.Field("a", T.int64, True), T.Field("b", T.int64, True)])
    bt = O.Batch(ins, [O.Col(T.int64, a, av), O.Col(T.int64, b, bv)], n)
    A, B = E.Column("a"), E.Column("b")
    pe = O.ProjectExec([(A, "a"), (E.BinaryExpr(A, "Plus", B), "c")], ins, [E.BinaryExpr(A, "Lt", E.Literal(500, T.int64))])
    out = O.concat_batches(pe.schema, pe.execute([bt.take(np.arange(i, min(i + 10000, n))) for i in range(0, n, 10000)]))
    assert np.array_equal(out.cols[0].values, oa) and np.array_equal(out.cols[1].valid, ocv)
    assert np.array_equal(out.cols[1].values[ocv], oc[ocv])


@pytest.mark.parametrize("nthreads", [1, 3])
@pytest.mark.parametrize("n", [0, 1, 35_000])
def test_q1_filter_agg_port_matches_numpy_oracle(n, nthreads):
    """M2 (q1 shape): Filter[f >= lo, f <= hi] -> SUM(v) GROUP BY k1, k2, Partial -> Final"""
    rng = np.random.default_rng(8)
    f = rng.integers(0, 1000, n, dtype=np.int64); k1 = rng.integers(0, 500, n, dtype=np.int64)
    k2 = rng.integers(0, 8, n, dtype=np.int64); v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    got = cpu_ref.q1_filter_agg(f, k1, k2, v, 200, 399, nthreads)
    ins = T.Schema([T.Field(c, T.int64, False) for c in ("f", "k1", "k2", "v")])
    ones = np.ones(n, bool)
    b = O.Batch(ins, [O.Col(T.int64, a, ones) for a in (f, k1, k2, v)], n)
    batches = [b.take(np.arange(i, min(i + 10000, n))) for i in range(0, n, 10000)]
    preds = [E.BinaryExpr(E.Column("f"), "GtEq", E.Literal(200, T.int64)), E.BinaryExpr(E.Column("f"), "LtEq", E.Literal(399, T.int64))]
    g = [E.GroupingExpr("k1", E.Column("k1")), E.GroupingExpr("k2", E.Column("k2"))]
    part = O.AggExec(E.HASH_AGG, g, [E.AggExpr("s", E.PARTIAL, E.AggFunctionExpr(E.AGG_SUM, [E.Column("v")], T.int64))], False, ins)
    fin = O.AggExec(E.HASH_AGG, g, [E.AggExpr("s", E.FINAL, E.AggFunctionExpr(E.AGG_SUM, [E.placeholder(T.int64)], T.int64))], False, part.schema)
    exp = O.rows_multiset(fin.execute(part.execute(O.FilterExec(preds, ins).execute(batches))))
    ms = {}
    for i in range(len(got["k1"])):
        key = (int(got["k1"][i]), int(got["k2"][i]), int(got["sum"][i]))
        ms[key] = ms.get(key, 0) + 1
    assert ms == exp
