"""Hash join on the GPU through the C ABI: every hash-join golden of the reference
(joins/test.rs:275-985) under its four hash drivers (BroadcastJoinExec / HashJoinExec nodes x map side left / right), and
seeded random inputs (NULL keys, duplicates on both sides, two keys, mixed widths, several batches) against the oracle
(oracle/join_oracle.py) for every join type — compared like the reference does, on the sorted rows."""
import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, native, plans as PL, types as T
from oracle import blaze_oracle as O
from oracle import join_oracle as J
from helpers import *
from join_goldens import CASES, arrow_batches

pytestmark = pytest.mark.gpu

# oracle join type -> wire JoinType (auron.proto:475-483)
WIRE = {J.INNER: PL.JOIN_INNER, J.LEFT: PL.JOIN_LEFT, J.RIGHT: PL.JOIN_RIGHT, J.FULL: PL.JOIN_FULL, J.LEFT_SEMI: PL.JOIN_SEMI, J.LEFT_ANTI: PL.JOIN_ANTI, J.EXISTENCE: PL.JOIN_EXISTENCE}


def _key(t):
    return tuple((x is None, x if x is not None else 0) for x in t)


def _rows(batches):
    rows = []
    for b in [O.batch_from_arrow(x) for x in batches]:
        for r in range(b.num_rows):
            rows.append(tuple(None if not c.valid[r] else (c.values[r].item() if hasattr(c.values[r], "item") else c.values[r]) for c in b.cols))
    return sorted(rows, key=_key)


def _run(lb, rb, on, jt, map_side, broadcast, conf=None):
    left, right = PL.MemoryExec.from_arrow(lb, lb[0].schema), PL.MemoryExec.from_arrow(rb, rb[0].schema)
    ls, rs = left.schema(), right.schema()
    schema = PL.build_join_schema(ls, rs, WIRE[jt])
    on_exprs = [(E.Column(l), E.Column(r)) for l, r in on]
    if broadcast:                                    # the map side arrives wrapped in BroadcastJoinBuildHashMapExec (joins/test.rs:211-244)
        if map_side == PL.LEFT_SIDE:
            left = PL.BroadcastJoinBuildHashMapExec(left, [l for l, _ in on_exprs])
        else:
            right = PL.BroadcastJoinBuildHashMapExec(right, [r for _, r in on_exprs])
    plan = PL.BroadcastJoinExec(schema, left, right, on_exprs, WIRE[jt], map_side, broadcast, "map-0" if broadcast else None)
    out = PL.collect(plan, conf)
    return plan, out


@pytest.mark.parametrize("driver", ["BHJLeftProbed", "BHJRightProbed", "SHJLeftProbed", "SHJRightProbed"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_join_goldens(case, driver):
    name, left, right, on, jt, expected = case[:6]
    dtype = case[6] if len(case) > 6 else "int32"
    lb, rb = arrow_batches(left, dtype), arrow_batches(right, dtype)
    map_side = PL.RIGHT_SIDE if "LeftProbed" in driver else PL.LEFT_SIDE
    plan, out = _run(lb, rb, on, jt, map_side, driver.startswith("BHJ"))
    conv = (lambda v: v) if dtype == "int32" else None
    got = _rows(out)
    if dtype != "int32":                              # date32 / timestamp cells come back as ints in the oracle's Col
        got = [tuple(None if x is None else int(x) for x in r) for r in got]
    assert got == sorted(expected, key=_key)
    assert plan.last_metrics["gpu_kernel_launches"] > 0 and plan.build_metrics["table_capacity_slots"] >= 1024


def _random_tables(seed, n_left, n_right, null_frac, two_keys):
    rng = np.random.default_rng(seed)
    def side(n, tag, krange):
        k = rng.integers(0, krange, n).astype(np.int32)
        k2 = rng.integers(0, 3, n, dtype=np.int64)
        cols = [with_nulls(rng, k, null_frac, pa.int32()), with_nulls(rng, k2, null_frac / 2), with_nulls(rng, rng.integers(-10**9, 10**9, n, dtype=np.int64), null_frac),
                pa.array(rng.normal(size=n)), pa.array(rng.integers(-100, 100, n).astype(np.int16), pa.int16())]
        names = [f"k{tag}", f"j{tag}", f"v{tag}", f"x{tag}", f"s{tag}"]
        return pa.RecordBatch.from_arrays(cols, names=names)
    return side(n_left, "l", 400), side(n_right, "r", 300)


@pytest.mark.parametrize("map_side", [PL.LEFT_SIDE, PL.RIGHT_SIDE])
@pytest.mark.parametrize("jt", [J.INNER, J.LEFT, J.RIGHT, J.FULL, J.LEFT_SEMI, J.LEFT_ANTI, J.EXISTENCE], ids=lambda j: J.NAMES[j])
@pytest.mark.parametrize("variant", ["one key", "two keys + nulls"])
def test_random_joins_match_the_oracle(jt, map_side, variant):
    two = variant.startswith("two")
    l, r = _random_tables(31 + jt, 6_000, 2_500, 0.1 if two else 0.0, two)
    lb, rb = split_batches(l, 1_700), split_batches(r, 900)
    on = [("kl", "kr")] + ([("jl", "jr")] if two else [])
    plan, out = _run(lb, rb, on, jt, map_side, False, native.default_conf(staging_rows=0))
    ls, rs = T.from_arrow_schema(l.schema), T.from_arrow_schema(r.schema)
    oj = J.HashJoin(ls, rs, [(l.schema.names.index(a), r.schema.names.index(b)) for a, b in on], jt, "left" if map_side == PL.LEFT_SIDE else "right")
    exp = oj.execute(oracle_batches(lb), oracle_batches(rb))
    assert O.rows_multiset([O.batch_from_arrow(b) for b in out]) == O.rows_multiset(exp)
    assert [f.name for f in oj.schema] == out[0].schema.names if out else True


def test_one_build_op_serves_several_probe_ops():
    """the map is built once and shared, like the process-wide cache keyed by cached_build_hash_map_id (broadcast_join_exec.rs:640-677)"""
    l, r = _random_tables(77, 4_000, 1_000, 0.05, False)
    ls, rs = T.from_arrow_schema(l.schema), T.from_arrow_schema(r.schema)
    build = PL.BroadcastJoinBuildHashMapExec(PL.MemoryExec.from_arrow([r], r.schema), [E.Column("kr")])
    with native.NativeOp(build.plan_bytes()) as bop:
        bop.push(r); bop.finish()
        for jt in (PL.JOIN_INNER, PL.JOIN_ANTI):
            halves = [l.slice(0, 2_000), l.slice(2_000, 2_000)]
            got = []
            for part in halves:                    # two "tasks" probe the same map
                plan = PL.BroadcastJoinExec(PL.build_join_schema(ls, rs, jt), PL.MemoryExec.from_arrow([part], l.schema), build, [(E.Column("kl"), E.Column("kr"))], jt, PL.RIGHT_SIDE, True, "m")
                with native.NativeOp(plan.plan_bytes()) as op:
                    op.attach_build(bop)
                    op.push(part); op.finish()
                    got += op.pull_all()
            oj = J.HashJoin(ls, rs, [(0, 0)], J.INNER if jt == PL.JOIN_INNER else J.LEFT_ANTI, "right")
            exp = oj.execute(oracle_batches([l]), oracle_batches([r]))
            assert O.rows_multiset([O.batch_from_arrow(b) for b in got]) == O.rows_multiset(exp)


def test_probe_without_a_build_side_is_a_state_error():
    l, r = _random_tables(5, 100, 100, 0.0, False)
    ls, rs = T.from_arrow_schema(l.schema), T.from_arrow_schema(r.schema)
    plan = PL.BroadcastJoinExec(PL.build_join_schema(ls, rs, PL.JOIN_INNER), PL.MemoryExec.from_arrow([l], l.schema), PL.MemoryExec.from_arrow([r], r.schema),
                                [(E.Column("kl"), E.Column("kr"))], PL.JOIN_INNER, PL.RIGHT_SIDE)
    with native.NativeOp(plan.plan_bytes(), native.default_conf(staging_rows=0)) as op:
        with pytest.raises(native.NativeError) as ei:
            op.push(l)
        assert ei.value.code == native.ERR_STATE


def test_empty_sides():
    l, r = _random_tables(6, 500, 200, 0.0, False)
    for lb, rb in (([l], [r.slice(0, 0)]), ([l.slice(0, 0)], [r])):
        for jt in (J.INNER, J.LEFT, J.FULL, J.LEFT_ANTI, J.EXISTENCE):
            for side in (PL.LEFT_SIDE, PL.RIGHT_SIDE):
                plan, out = _run(lb, rb, [("kl", "kr")], jt, side, False, native.default_conf(staging_rows=0))
                oj = J.HashJoin(T.from_arrow_schema(l.schema), T.from_arrow_schema(r.schema), [(0, 0)], jt, "left" if side == PL.LEFT_SIDE else "right")
                exp = oj.execute(oracle_batches(lb), oracle_batches(rb))
                assert O.rows_multiset([O.batch_from_arrow(b) for b in out]) == O.rows_multiset(exp), (J.NAMES[jt], side)
