"""The reference's own cast / Spark-decimal known-answer vectors (tests/kat_cases.py) through the CUDA path and the
C ABI: ProjectExec[expr] over a one-column batch, compared with the EXPECTED column of the reference test itself (not
with the oracle), once as a bare projection and once below a never-false filter (the fused filter+project kernel)."""
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, plans as PL, types as T, native
import kat_cases as K

pytestmark = pytest.mark.gpu

CASES = K.cases()


@pytest.mark.parametrize("with_filter", [False, True])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_kat_through_the_c_abi(case, with_filter):
    name, ref, inp, expr, exp = case
    rb = pa.RecordBatch.from_arrays([inp, pa.array(range(len(inp)), pa.int64())], names=["x", "i"])
    plan = PL.MemoryExec.from_arrow([rb], rb.schema)
    if with_filter:
        plan = PL.FilterExec([E.BinaryExpr(E.Column("i"), "GtEq", E.Literal(0, T.int64))], plan)
    plan = PL.ProjectExec([(expr, "y")], plan)
    got = PL.collect(plan, native.default_conf(staging_rows=0))
    col = pa.concat_arrays([b.column(0) for b in got]) if got else pa.array([], exp.type)
    assert K.same_column(col, exp), f"{name} ({ref}): GPU gives {col.to_pylist()}, the reference test expects {exp.to_pylist()}"
    assert plan.last_metrics["gpu_kernel_launches"] > 0
