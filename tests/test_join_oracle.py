"""oracle/join_oracle.py against every hash-join golden of the reference (joins/test.rs:275-985) under both probe sides."""
import pytest

from blaze_b200 import types as T
from oracle import blaze_oracle as O
from oracle import join_oracle as J
from join_goldens import CASES, arrow_batches


def sorted_rows(batches):
    rows = []
    for b in batches:
        for r in range(b.num_rows):
            rows.append(tuple(None if not c.valid[r] else (c.values[r].item() if hasattr(c.values[r], "item") else c.values[r]) for c in b.cols))
    return sorted(rows, key=lambda t: tuple((x is None, x if x is not None else 0) for x in t))


@pytest.mark.parametrize("build_side", ["right", "left"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_join_goldens(case, build_side):
    name, left, right, on, jt, expected = case[:6]
    dtype = case[6] if len(case) > 6 else "int32"
    lb, rb = arrow_batches(left, dtype), arrow_batches(right, dtype)
    ls, rs = T.from_arrow_schema(lb[0].schema), T.from_arrow_schema(rb[0].schema)
    on_idx = [(lb[0].schema.names.index(l), rb[0].schema.names.index(r)) for l, r in on]
    j = J.HashJoin(ls, rs, on_idx, jt, build_side)
    out = j.execute([O.batch_from_arrow(b) for b in lb], [O.batch_from_arrow(b) for b in rb])
    exp = sorted(expected, key=lambda t: tuple((x is None, x if x is not None else 0) for x in t))
    assert sorted_rows(out) == exp
    assert len(j.schema) == len(expected[0])


def test_schema_rules():
    ls = T.Schema([T.Field("a", T.int32, False)]); rs = T.Schema([T.Field("b", T.int64, False)])
    assert [f.nullable for f in J.join_schema(ls, rs, J.LEFT)] == [False, True]
    assert [f.nullable for f in J.join_schema(ls, rs, J.RIGHT)] == [True, False]
    assert [f.nullable for f in J.join_schema(ls, rs, J.FULL)] == [True, True]
    assert [f.name for f in J.join_schema(ls, rs, J.EXISTENCE)] == ["a", "exists#0"]
    assert [f.name for f in J.join_schema(ls, rs, J.RIGHT_SEMI)] == ["b"]
