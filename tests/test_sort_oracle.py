"""oracle/sort_oracle.py against the reference's golden (sort_exec.rs:1447-1476) and its fuzz property (:1527-1607)."""
import numpy as np
import pyarrow as pa

from oracle import blaze_oracle as O
from oracle import sort_oracle as S


def test_reference_golden_sort_i32_with_fetch():
    rb = pa.RecordBatch.from_arrays([pa.array(x, pa.int32()) for x in ([9, 8, 7, 6, 5, 4, 3, 2, 1, 0], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9], [5, 6, 7, 8, 9, 0, 1, 2, 3, 4])], names=["a", "b", "c"])
    out = S.sort_exec([O.batch_from_arrow(rb)], [(0, False, True)], fetch=6)
    rows = [tuple(int(c.values[r]) for c in out.cols) for r in range(out.num_rows)]
    assert rows == [(0, 9, 4), (1, 8, 3), (2, 7, 2), (3, 6, 1), (4, 5, 0), (5, 4, 9)]


def test_order_matches_an_independent_engine():
    """the fuzz test's property: same rows as DataFusion's own SortExec — here pyarrow's sort_indices plays the second engine"""
    rng = np.random.default_rng(2)
    n = 5000
    k1 = pa.array(rng.integers(-50, 50, n), pa.int64(), mask=rng.random(n) < 0.1)
    k2 = pa.array(rng.normal(size=n), pa.float64(), mask=rng.random(n) < 0.1)
    v = pa.array(np.arange(n), pa.int32())
    rb = pa.RecordBatch.from_arrays([k1, k2, v], names=["k1", "k2", "v"])
    for desc1, nf in ((False, True), (True, False)):
        out = S.sort_exec([O.batch_from_arrow(rb.slice(0, 3000)), O.batch_from_arrow(rb.slice(3000))], [(0, desc1, nf), (1, not desc1, nf)])
        idx = pa.compute.sort_indices(pa.Table.from_batches([rb]), sort_keys=[("k1", "descending" if desc1 else "ascending"), ("k2", "ascending" if desc1 else "descending")],
                                      null_placement="at_start" if nf else "at_end")
        exp = rb.take(idx)
        got_keys = [(None if not out.cols[0].valid[r] else int(out.cols[0].values[r]), None if not out.cols[1].valid[r] else float(out.cols[1].values[r])) for r in range(n)]
        exp_keys = list(zip(exp.column(0).to_pylist(), exp.column(1).to_pylist()))
        assert got_keys == exp_keys
        assert sorted(int(x) for x in out.cols[2].values) == list(range(n))


def test_float_total_order_and_nulls():
    vals = [0.0, -0.0, float("inf"), float("-inf"), float("nan"), 1.5, None, -2.0]
    rb = pa.RecordBatch.from_arrays([pa.array(vals, pa.float64())], names=["x"])
    out = S.sort_exec([O.batch_from_arrow(rb)], [(0, False, False)])
    got = [None if not out.cols[0].valid[r] else float(out.cols[0].values[r]) for r in range(8)]
    assert got[:4] == [float("-inf"), -2.0, -0.0, 0.0] and str(got[2]) == "-0.0" and got[4:6] == [1.5, float("inf")] and got[6] != got[6] and got[7] is None
