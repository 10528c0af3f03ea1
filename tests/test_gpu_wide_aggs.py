"""The WIDE tile aggregates (kernels_tile.cu): f64 SUM / AVG, decimal128 SUM / AVG, integer and f64 MIN / MAX over one or
two dense integer keys, with NULL arguments, NULL keys, fused conjuncts and later batches that leave the dense key range
(hashed fall-back) — Partial -> Final through the C ABI vs the oracle; the same plans on the generic bytecode kernel agree."""
import decimal

import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, plans as PL, types as T, native
from oracle import blaze_oracle as O
from helpers import *

pytestmark = pytest.mark.gpu

D172 = pa.decimal128(17, 2)


def _table(n, seed, null_frac, key_null_frac, two_keys):
    rng = np.random.default_rng(seed)
    k = rng.integers(-3, 700, n).astype(np.int32)
    k[n * 2 // 3:] += rng.integers(0, 5000, n - n * 2 // 3).astype(np.int32)          # later rows leave the dense range of the first batch
    k2 = rng.integers(0, 6, n, dtype=np.int64)
    f = rng.integers(0, 100, n, dtype=np.int64)
    x = rng.normal(0, 1e6, n); x[rng.random(n) < 0.01] = -0.0
    i = rng.integers(-2**40, 2**40, n, dtype=np.int64)
    raw = rng.integers(-10**15, 10**15, n)
    d = pa.array([decimal.Decimal(int(r)).scaleb(-2) for r in raw], type=D172, mask=(rng.random(n) < null_frac) if null_frac else None)
    cols = {"k": with_nulls(rng, k, key_null_frac, pa.int32()), "k2": pa.array(k2), "f": pa.array(f), "x": with_nulls(rng, x, null_frac), "i": with_nulls(rng, i, null_frac), "d": d}
    names = ["k", "k2", "f", "x", "i", "d"]
    return pa.RecordBatch.from_arrays([cols[c] for c in names], names=names)


SHAPES = {
    "sum f64 + count":        ([("s", E.AGG_SUM, "x", T.float64), ("c", E.AGG_COUNT, "x", T.int64)], (0,)),
    "avg f64":                ([("a", E.AGG_AVG, "x", T.float64)], (0,)),
    "avg int -> f64":         ([("a", E.AGG_AVG, "i", T.float64), ("c", E.AGG_COUNT, "i", T.int64)], (0,)),
    "sum dec + count":        ([("s", E.AGG_SUM, "d", T.decimal128(27, 2)), ("c", E.AGG_COUNT, "d", T.int64)], ()),
    "avg dec":                ([("a", E.AGG_AVG, "d", T.decimal128(21, 6))], ()),
    "min max int":            ([("mn", E.AGG_MIN, "i", T.int64), ("mx", E.AGG_MAX, "i", T.int64)], ()),
    "min max f64":            ([("mn", E.AGG_MIN, "x", T.float64), ("mx", E.AGG_MAX, "x", T.float64)], ()),
    "min max of two columns": ([("mn", E.AGG_MIN, "i", T.int64), ("mx", E.AGG_MAX, "i", T.int64), ("mf", E.AGG_MIN, "f", T.int64), ("xf", E.AGG_MAX, "f", T.int64)], ()),
    "sum int + sum int + count*": ([("s1", E.AGG_SUM, "i", T.int64), ("s2", E.AGG_SUM, "f", T.int64), ("n", E.AGG_COUNT, None, T.int64)], ()),
}


@pytest.mark.parametrize("variant", ["plain", "nulls", "two keys + filter + null keys"])
@pytest.mark.parametrize("shape", list(SHAPES))
def test_wide_tile_aggregates(shape, variant):
    specs, fcols = SHAPES[shape]
    two = variant.startswith("two")
    rb = _table(60_000, 11, 0.0 if variant == "plain" else 0.15, 0.02 if two else 0.0, two)
    batches = split_batches(rb, 20_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    groupings = [E.GroupingExpr("k", E.Column("k"))] + ([E.GroupingExpr("k2", E.Column("k2"))] if two else [])
    preds = [E.BinaryExpr(E.Column("f"), "GtEq", E.Literal(10, T.int64)), E.BinaryExpr(E.Column("f"), "Lt", E.Literal(85, T.int64))] if two else []
    ch = lambda col: [E.Column(col)] if col is not None else [E.Literal(1, T.int64)]
    mk = lambda mode, src: [E.AggExpr(nm, mode, PL.create_agg(fn, ch(col) if mode == E.PARTIAL else [E.placeholder(rt)], src, rt)) for nm, fn, col, rt in specs]
    nk = len(groupings)
    outs = {}
    for mode, conf in (("wide", native.default_conf(staging_rows=0)), ("generic", native.default_conf(staging_rows=0, force_generic_kernels=1))):
        partial = PL.AggExec(PL.HashAgg, groupings, mk(E.PARTIAL, ins), False, PL.FilterExec(preds, leaf) if preds else leaf)
        final = PL.AggExec(PL.HashAgg, groupings, mk(E.FINAL, partial.schema()), False, partial)
        outs[mode] = PL.collect(final, conf)
        if mode == "wide":
            assert final.last_metrics["fast_path_launches"] > 0, "the plan must take the wide tile kernel"
    ob = oracle_batches(batches)
    op = O.AggExec(E.HASH_AGG, groupings, mk(E.PARTIAL, ins), False, ins)
    of = O.AggExec(E.HASH_AGG, groupings, mk(E.FINAL, op.schema), False, op.schema)
    exp = of.execute(op.execute(O.FilterExec(preds, ins).execute(ob) if preds else ob))
    fc = tuple(nk + c for c in fcols)
    assert_multiset_equal(outs["wide"], exp, fc)
    assert_multiset_equal(outs["generic"], exp, fc)


def test_wide_partial_state_is_byte_exact():
    """the frozen Binary state column of a Partial stage that ran on the wide kernel (decimal SUM + COUNT: exact integers)"""
    rb = _table(30_000, 12, 0.1, 0.0, False)
    batches = split_batches(rb, 10_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    g = [E.GroupingExpr("k", E.Column("k"))]
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("d")], ins, T.decimal128(27, 2))),
            E.AggExpr("c", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Column("d")], ins, T.int64))]
    plan = PL.AggExec(PL.HashAgg, g, aggs, False, leaf)
    got = PL.collect(plan, native.default_conf(staging_rows=0))
    assert plan.last_metrics["fast_path_launches"] > 0
    exp = O.AggExec(E.HASH_AGG, g, aggs, False, ins).execute(oracle_batches(batches))
    assert_multiset_equal(got, exp)
