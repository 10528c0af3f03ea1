"""SortExec on the GPU through the C ABI: the reference's golden (sort_exec.rs:1447-1476) and seeded
inputs against the oracle — the key columns must come out in exactly the oracle's order (the order among equal keys is
unspecified in the reference, so full rows are compared as multisets) and `fetch` keeps the first rows."""
import decimal

import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, native, plans as PL, types as T
from oracle import blaze_oracle as O
from oracle import sort_oracle as S
from helpers import *

pytestmark = pytest.mark.gpu


def _keys_of(batch, exprs):
    return [S.sort_key(batch.cols, r, exprs) for r in range(batch.num_rows)]


def _run(batches, exprs, fetch=None, conf=None):
    leaf = PL.MemoryExec.from_arrow(batches, batches[0].schema)
    names = batches[0].schema.names
    plan = PL.SortExec(leaf, [(E.Column(names[c]), d, nf) for c, d, nf in exprs], fetch)
    out = PL.collect(plan, conf)
    schema = T.from_arrow_schema(batches[0].schema)
    got = O.concat_batches(schema, [O.batch_from_arrow(b) for b in out])
    return plan, got


def test_reference_golden_sort_i32_with_fetch():
    rb = pa.RecordBatch.from_arrays([pa.array(x, pa.int32()) for x in ([9, 8, 7, 6, 5, 4, 3, 2, 1, 0], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9], [5, 6, 7, 8, 9, 0, 1, 2, 3, 4])],
                                    schema=pa.schema([pa.field(n, pa.int32(), False) for n in "abc"]))
    plan, got = _run([rb], [(0, False, True)], fetch=6)
    assert [tuple(int(c.values[r]) for c in got.cols) for r in range(got.num_rows)] == [(0, 9, 4), (1, 8, 3), (2, 7, 2), (3, 6, 1), (4, 5, 0), (5, 4, 9)]
    assert "SortExec [a@0 ASC NULLS FIRST] fetch=6" in plan.explain()


def _table(n, seed, null_frac):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 100, n); x[rng.random(n) < 0.02] = -0.0; x[rng.random(n) < 0.01] = np.inf; x[rng.random(n) < 0.01] = np.nan
    cols = {
        "i64": with_nulls(rng, rng.integers(-2**62, 2**62, n, dtype=np.int64), null_frac),
        "i32": with_nulls(rng, rng.integers(-1000, 1000, n).astype(np.int32), null_frac, pa.int32()),
        "i8": with_nulls(rng, rng.integers(-128, 128, n).astype(np.int8), null_frac, pa.int8()),
        "f64": with_nulls(rng, x, null_frac),
        "f32": with_nulls(rng, rng.normal(size=n).astype(np.float32), null_frac, pa.float32()),
        "d": pa.array(rng.integers(0, 300, n).astype(np.int32), pa.int32()).cast(pa.date32()),
        "dec": pa.array([decimal.Decimal(int(v)).scaleb(-2) for v in rng.integers(-10**17, 10**17, n)], pa.decimal128(30, 2), mask=(rng.random(n) < null_frac) if null_frac else None),
        "row": pa.array(np.arange(n, dtype=np.int64)),
    }
    return pa.RecordBatch.from_arrays(list(cols.values()), names=list(cols))


CASES = {
    "int64 asc": [(0, False, True)],
    "int32 desc nulls last, int8 asc": [(1, True, False), (2, False, True)],
    "f64 asc nulls last (totalOrder)": [(3, False, False)],
    "date desc, f32 desc nulls first, int64 asc": [(5, True, True), (4, True, True), (0, False, False)],
    "decimal128 desc nulls first": [(6, True, True)],
    "int8 asc, decimal asc nulls last": [(2, False, True), (6, False, False)],
}


@pytest.mark.parametrize("null_frac", [0.0, 0.15])
@pytest.mark.parametrize("case", list(CASES))
def test_sorted_like_the_oracle(case, null_frac):
    exprs = CASES[case]
    rb = _table(20_000, 3, null_frac)
    batches = split_batches(rb, 6_000)
    plan, got = _run(batches, exprs, conf=native.default_conf(staging_rows=0))
    exp = S.sort_exec(oracle_batches(batches), exprs)
    assert got.num_rows == exp.num_rows == rb.num_rows
    assert _keys_of(got, exprs) == _keys_of(exp, exprs)                 # the key columns come out in the oracle's order
    assert O.rows_multiset([got]) == O.rows_multiset([exp])             # and no row is lost or altered
    assert plan.last_metrics["gpu_kernel_launches"] > 0


def test_ties_keep_arrival_order():
    """the radix passes are stable: with few distinct keys the `row` column is ascending inside every key"""
    rb = _table(30_000, 9, 0.0)
    _, got = _run(split_batches(rb, 10_000), [(2, False, True)], conf=native.default_conf(staging_rows=0))
    k, row = got.cols[2].values, got.cols[7].values
    assert (np.diff(k.astype(np.int64)) >= 0).all()
    same = np.diff(k.astype(np.int64)) == 0
    assert (np.diff(row)[same] > 0).all()


@pytest.mark.parametrize("fetch", [0, 1, 100, 10**6])
def test_fetch_keeps_the_first_rows(fetch):
    rb = _table(5_000, 4, 0.1)
    exprs = [(0, True, False)]
    plan, got = _run([rb], exprs, fetch=fetch)
    exp = S.sort_exec(oracle_batches([rb]), exprs, fetch)
    assert got.num_rows == min(fetch, 5_000)
    if fetch:
        assert _keys_of(got, exprs) == _keys_of(exp, exprs)


def test_sort_after_an_aggregate_and_empty_input():
    """ORDER BY on the Final aggregate's output (the tail of q1 / q3): AggExec(Partial) -> AggExec(Final) -> SortExec fetch 10, one op"""
    rng = np.random.default_rng(5)
    n = 50_000
    rb = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 2000, n, dtype=np.int64)), pa.array(rng.integers(-10**6, 10**6, n, dtype=np.int64))], names=["k", "v"])
    leaf = PL.MemoryExec.from_arrow(split_batches(rb, 10_000), rb.schema)
    ins = leaf.schema()
    g = [E.GroupingExpr("k", E.Column("k"))]
    mk = lambda mode, ch: [E.AggExpr("s", mode, PL.create_agg(E.AGG_SUM, ch, ins, T.int64))]
    partial = PL.AggExec(PL.HashAgg, g, mk(E.PARTIAL, [E.Column("v")]), False, leaf)
    final = PL.AggExec(PL.HashAgg, g, mk(E.FINAL, [E.placeholder(T.int64)]), False, partial)
    plan = PL.SortExec(final, [(E.Column("s"), True, False), (E.Column("k"), False, True)], 10)
    out = PL.collect(plan)
    sums = {}
    for kk, vv in zip(rb.column(0).to_numpy(), rb.column(1).to_numpy()):
        sums[int(kk)] = sums.get(int(kk), 0) + int(vv)
    exp = sorted(sums.items(), key=lambda kv: (-kv[1], kv[0]))[:10]
    got = [(int(a), int(b)) for b_ in out for a, b in zip(b_.column(0).to_pylist(), b_.column(1).to_pylist())]
    assert got == exp
    empty = PL.SortExec(PL.MemoryExec.from_arrow([rb.slice(0, 0)], rb.schema), [(E.Column("k"), False, True)])
    assert PL.collect(empty) == []
