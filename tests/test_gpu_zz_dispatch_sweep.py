"""Sweep of the HashAgg kernel dispatch (kernels_fast.cu: launch_agg_fast_update): every combination of
{1,2 keys} x {tiny, mid, sparse key ranges} x {non-null int64, typed + NULLs} x {aggregate sets} x {fused filter or not}
against the oracle.  Tiny ranges take the shared-memory tables, mid ranges the dense global table (gang or
one-row-per-lane form, 2- or 4-word entries), sparse ranges the hashed kernels.  (Named zz: added late in round 1 —
it runs after the files whose cases were each validated on hardware individually.)"""
import itertools

import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, plans as PL, types as T, native
from oracle import blaze_oracle as O
from helpers import *

pytestmark = pytest.mark.gpu

AGG_SETS = {
    "sum": lambda ins: [("s", E.AGG_SUM, [E.Column("v")])],
    "sum+count(v)": lambda ins: [("s", E.AGG_SUM, [E.Column("v")]), ("c", E.AGG_COUNT, [E.Column("v")])],
    "sum+count(*)": lambda ins: [("s", E.AGG_SUM, [E.Column("v")]), ("n", E.AGG_COUNT, [E.Literal(1, T.int64)])],
    "count(*)": lambda ins: [("n", E.AGG_COUNT, [E.Literal(1, T.int64)])],
    "sum(v)+sum(w)": lambda ins: [("s", E.AGG_SUM, [E.Column("v")]), ("t", E.AGG_SUM, [E.Column("w")])],
}
RANGES = {"tiny": (12, 3), "mid": (9000, 6), "sparse": (2**40, 2**20)}


@pytest.mark.parametrize("nkeys,rng_name,typed,aggset,filt", list(itertools.product([1, 2], RANGES, [False, True], AGG_SETS, [False, True])))
def test_dispatch_combination(nkeys, rng_name, typed, aggset, filt):
    n = 60_000
    rng = np.random.default_rng(hash((nkeys, rng_name, typed, aggset, filt)) % 2**32)
    r0, r1 = RANGES[rng_name]
    k0 = rng.integers(-5, -5 + r0, n, dtype=np.int64)
    k1 = rng.integers(100, 100 + r1, n, dtype=np.int64)
    if rng_name == "sparse":                                       # few distinct values spread over a huge range
        k0 = rng.choice(rng.integers(-2**40, 2**40, 700, dtype=np.int64), n)
        k1 = rng.choice(rng.integers(0, 2**20, 9, dtype=np.int64), n)
    k0[40_000:] += rng.integers(0, 3, n - 40_000) * (r0 if rng_name != "sparse" else 1)      # later batches step outside the first batch's range
    v = rng.integers(-2**45, 2**45, n, dtype=np.int64)
    w = rng.integers(-1000, 1000, n, dtype=np.int64)
    f = rng.integers(0, 10, n, dtype=np.int64)
    if typed:
        small = rng_name != "sparse"
        cols = [with_nulls(rng, k0.astype(np.int32) if small else k0, 0.02, pa.int32() if small else pa.int64()),
                with_nulls(rng, k1.astype(np.int16) if small else k1, 0.03, pa.int16() if small else pa.int64()),
                with_nulls(rng, v, 0.15), with_nulls(rng, w.astype(np.int32), 0.1, pa.int32()), with_nulls(rng, f.astype(np.int8), 0.05, pa.int8())]
        rb = rb_from_cols(["k0", "k1", "v", "w", "f"], cols)
    else:
        rb = rb_from_cols(["k0", "k1", "v", "w", "f"], [pa.array(x) for x in (k0, k1, v, w, f)])
    batches = split_batches(rb, 20_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    groupings = [E.GroupingExpr("k0", E.Column("k0"))] + ([E.GroupingExpr("k1", E.Column("k1"))] if nkeys == 2 else [])
    aggs = [E.AggExpr(nm, E.PARTIAL, PL.create_agg(fn, ch, ins, T.int64)) for nm, fn, ch in AGG_SETS[aggset](ins)]
    ftype = T.int8 if typed else T.int64
    preds = [E.BinaryExpr(E.Column("f"), "GtEq", E.Literal(2, ftype)), E.BinaryExpr(E.Column("f"), "NotEq", E.Literal(7, ftype))] if filt else []
    child = PL.FilterExec(preds, leaf) if preds else leaf
    plan = PL.AggExec(PL.HashAgg, groupings, aggs, False, child)
    got = PL.collect(plan, native.default_conf(staging_rows=0))
    ob = oracle_batches(batches)
    exp = O.AggExec(E.HASH_AGG, groupings, aggs, False, ins).execute(O.FilterExec(preds, ins).execute(ob) if preds else ob)
    assert_multiset_equal(got, exp)
    assert plan.last_metrics["fast_path_launches"] > 0
