"""Seeded golden cases of the hot path (inputs + the plans run on them).

The reference (Rust, un-buildable here — DESIGN.md §0) has no exportable vectors for these shapes beyond the KATs
pinned in tests/test_oracle_golden.py, so the expected outputs are produced by the pinned numpy oracle ONCE, by
tests/golden/make_golden.py, and committed as Arrow IPC files.  The CPU suite checks that the oracle still
reproduces them (drift guard); the GPU suite compares the CUDA path against the files without running the oracle.
"""
import decimal

import numpy as np
import pyarrow as pa

from blaze_b200 import exprs as E, plans as PL, types as T


def _nulls(rng, values, frac, pa_type=None):
    if frac <= 0:
        return pa.array(values, type=pa_type)
    return pa.array(values, mask=rng.random(len(values)) < frac, type=pa_type)


def _rb(names, arrays):
    return pa.RecordBatch.from_arrays(arrays, names=names)


class Case:
    """name, input batch, rows per pushed batch, `build(leaf) -> plan` (FilterExec/ProjectExec/AggExec over `leaf`),
    ordered (Filter/Project) or multiset (Agg) comparison, float columns compared at 1e-6"""
    def __init__(self, name, rb, batch_rows, build, ordered, float_cols=()):
        self.name, self.rb, self.batch_rows, self.build, self.ordered, self.float_cols = name, rb, batch_rows, build, ordered, float_cols


def _m0():
    rng = np.random.default_rng(1001)
    n = 5003
    rb = _rb(["a", "b"], [_nulls(rng, rng.integers(0, 1000, n, dtype=np.int64), 0.1), _nulls(rng, rng.integers(-2**31, 2**31, n, dtype=np.int64), 0.1)])
    A, B = E.Column("a"), E.Column("b")
    return Case("m0_filter_project", rb, 1000, lambda leaf: PL.ProjectExec([(A, "a"), (E.BinaryExpr(A, "Plus", B), "c")],
                PL.FilterExec([E.BinaryExpr(A, "Lt", E.Literal(500, T.int64))], leaf)), True)


def _m1(mode_name):
    rng = np.random.default_rng(1002)
    n = 3001
    rb = _rb(["k", "v"], [_nulls(rng, rng.integers(-20, 30, n, dtype=np.int64), 0.03), _nulls(rng, rng.integers(-2**62, 2**62, n, dtype=np.int64), 0.25)])

    def build(leaf):
        ins = leaf.schema()
        g = [E.GroupingExpr("k", E.Column("k"))]
        specs = [("s", E.AGG_SUM, [E.Column("v")], T.int64), ("c", E.AGG_COUNT, [E.Column("v")], T.int64)]
        partial = PL.AggExec(PL.HashAgg, g, [E.AggExpr(nm, E.PARTIAL, PL.create_agg(fn, ch, ins, rt)) for nm, fn, ch, rt in specs], False, leaf)
        if mode_name == "partial":
            return partial                                       # output carries the frozen Binary state column, byte for byte
        fin = [E.AggExpr(nm, E.FINAL, PL.create_agg(fn, [E.placeholder(ch[0].data_type(ins))], partial.schema(), rt)) for nm, fn, ch, rt in specs]
        return PL.AggExec(PL.HashAgg, g, fin, False, partial)
    return Case("m1_sum_count_" + mode_name, rb, 700, build, False)


def _q1():
    rng = np.random.default_rng(1003)
    n = 8000
    d = pa.array(rng.integers(0, 3000, n, dtype=np.int32), pa.int32()).cast(pa.date32())
    rb = _rb(["d", "k1", "k2", "v"], [d, _nulls(rng, rng.integers(0, 40, n).astype(np.int16), 0.02, pa.int16()), pa.array(rng.integers(-3, 3, n).astype(np.int8)),
                                      _nulls(rng, rng.integers(-1000, 1000, n).astype(np.int32), 0.1, pa.int32())])

    def build(leaf):
        ins = leaf.schema()
        preds = [E.BinaryExpr(E.Column("d"), "GtEq", E.Literal(1000, T.date32)), E.BinaryExpr(E.Literal(2000, T.date32), "Gt", E.Column("d")),
                 E.BinaryExpr(E.Column("k2"), "NotEq", E.Literal(0, T.int8))]
        g = [E.GroupingExpr("k1", E.Column("k1")), E.GroupingExpr("k2", E.Column("k2"))]
        aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], ins, T.int64)),
                E.AggExpr("n", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Literal(1, T.int64)], ins, T.int64))]
        return PL.AggExec(PL.HashAgg, g, aggs, True, PL.FilterExec(preds, leaf))
    return Case("q1_two_keys_fused_filter_partial", rb, 1000, build, False)


def _dec():
    rng = np.random.default_rng(1004)
    n = 4000
    raw = rng.integers(-10**15, 10**15, n)
    dec = pa.array([decimal.Decimal(int(r)).scaleb(-2) for r in raw], type=pa.decimal128(17, 2))
    rb = _rb(["k", "x", "d"], [pa.array(rng.integers(0, 60, n, dtype=np.int64)), _nulls(rng, rng.normal(0, 1e6, n), 0.1), dec])

    def build(leaf):
        ins = leaf.schema()
        g = [E.GroupingExpr("k", E.Column("k"))]
        specs = [("sx", E.AGG_SUM, [E.Column("x")], T.float64), ("ax", E.AGG_AVG, [E.Column("x")], T.float64),
                 ("mnx", E.AGG_MIN, [E.Column("x")], T.float64), ("mxx", E.AGG_MAX, [E.Column("x")], T.float64),
                 ("sd", E.AGG_SUM, [E.Column("d")], T.decimal128(27, 2)), ("ad", E.AGG_AVG, [E.Column("d")], T.decimal128(21, 6)),
                 ("mnd", E.AGG_MIN, [E.Column("d")], T.decimal128(17, 2)), ("mxd", E.AGG_MAX, [E.Column("d")], T.decimal128(17, 2)),
                 ("c", E.AGG_COUNT, [E.Column("x")], T.int64)]
        partial = PL.AggExec(PL.HashAgg, g, [E.AggExpr(nm, E.PARTIAL, PL.create_agg(fn, ch, ins, rt)) for nm, fn, ch, rt in specs], False, leaf)
        fin = [E.AggExpr(nm, E.FINAL, PL.create_agg(fn, [E.placeholder(ch[0].data_type(ins))], partial.schema(), rt)) for nm, fn, ch, rt in specs]
        return PL.AggExec(PL.HashAgg, g, fin, False, partial)
    return Case("f64_decimal_all_aggs_final", rb, 900, build, False, float_cols=(1, 2))


def all_cases():
    return [_m0(), _m1("partial"), _m1("final"), _q1(), _dec()]


def murmur3_case():
    """key columns (int64, nullable int32) and the partition counts of the committed partition-id vectors"""
    rng = np.random.default_rng(1005)
    n = 2000
    a = rng.integers(-2**62, 2**62, n, dtype=np.int64); a[:5] = [1, 0, -1, 2**63 - 1, -2**63]
    b = rng.integers(-2**31, 2**31, n, dtype=np.int64).astype(np.int32)
    return a, b, rng.random(n) >= 0.1, (2, 7, 200)
