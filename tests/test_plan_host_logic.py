"""Host logic without a GPU: plan decoding (hand-written proto3 + Arrow-IPC readers), create_agg
rewrites, schema/type inference, validation errors — through b200q_plan_explain."""
import pytest

from blaze_b200 import exprs as E, native, plans as PL, proto as P, types as T

S = T.Schema([T.Field("a", T.int64, False), T.Field("b", T.int64, True), T.Field("d", T.decimal128(12, 2), True),
              T.Field("f", T.float64, True), T.Field("i", T.int32, True), T.Field("t", T.date32, False)])
A, B, D, F, I, TT = (E.Column(c) for c in "abdfit")


def test_explain_filter_project_agg():
    leaf = PL.MemoryExec(S)
    f = PL.FilterExec([E.BinaryExpr(A, "Lt", E.Literal(500, T.int64)), E.IsNotNull(B)], leaf)
    p = PL.ProjectExec([(A, "a"), (E.BinaryExpr(A, "Plus", B), "c"), (E.TryCast(I, T.int64), "i64")], f)
    txt = p.explain()
    assert "ProjectExec [a@0 AS a, (a@0 Plus b@1) AS c, TryCast(i@4 AS int64) AS i64] schema=[a:int64, c:int64?, i64:int64?]" in txt
    assert "FilterExec [(a@0 Lt 500:int64), IsNotNull(b@1)]" in txt and "FFIReader schema=[a:int64, b:int64?" in txt
    assert [f.name for f in p.schema()] == ["a", "c", "i64"] and [f.nullable for f in p.schema()] == [False, True, True]


def test_create_agg_rewrites():
    leaf = PL.MemoryExec(S)
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [I], S, T.int64)),
            E.AggExpr("c1", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [A], S, T.int64)),          # non-nullable child dropped -> COUNT(*)
            E.AggExpr("c2", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [B, E.Literal(1, T.int64)], S, T.int64)),
            E.AggExpr("m", E.PARTIAL, PL.create_agg(E.AGG_MAX, [D], S, T.int32)),              # Max takes the child's type
            E.AggExpr("av", E.PARTIAL, PL.create_agg(E.AGG_AVG, [D], S, T.decimal128(16, 6)))]
    txt = PL.AggExec(PL.HashAgg, [E.GroupingExpr("g", TT)], aggs, True, leaf).explain()
    assert "Sum(TryCast(i@4 AS int64)):int64/Partial AS s" in txt            # agg.rs:190-193
    assert "Count():int64/Partial AS c1" in txt and "Count(b@1):int64/Partial AS c2" in txt   # agg.rs:178-189
    assert "Max(d@2):decimal128(12,2)/Partial AS m" in txt                   # agg.rs:198-201
    assert "Avg(TryCast(d@2 AS decimal128(16,6))):decimal128(16,6)/Partial AS av" in txt
    assert "schema=[g:date32, #9223372036854775807:binary]" in txt and "partial_skipping=true" in txt


def test_final_schema_and_modes():
    leaf = PL.MemoryExec(T.Schema([T.Field("k", T.int64, True), T.Field(E.AGG_BUF_COLUMN_NAME, T.binary, False)]))
    aggs = [E.AggExpr("s", E.FINAL, PL.create_agg(E.AGG_SUM, [E.placeholder(T.int64)], S, T.int64)),
            E.AggExpr("c", E.FINAL, PL.create_agg(E.AGG_COUNT, [E.placeholder(T.int64)], S, T.int64)),
            E.AggExpr("av", E.FINAL, PL.create_agg(E.AGG_AVG, [E.placeholder(T.float64)], S, T.float64))]
    plan = PL.AggExec(PL.HashAgg, [E.GroupingExpr("k", E.Column("k"))], aggs, False, leaf)
    assert "schema=[k:int64?, s:int64?, c:int64, av:float64?]" in plan.explain()
    mixed = aggs[:1] + [E.AggExpr("p", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Column("k")], S, T.int64))]
    with pytest.raises(native.NativeError, match="final aggregates may not exist along with partial"):   # agg_ctx.rs:115
        PL.AggExec(PL.HashAgg, [], mixed, False, leaf)


@pytest.mark.parametrize("value,dt", [(True, T.bool_), (-5, T.int8), (300, T.int16), (-70000, T.int32), (2**40, T.int64), (1.5, T.float32),
                                      (-2.25, T.float64), (12345, T.date32), (10**15, T.timestamp_us), (-123456789, T.decimal128(20, 4)),
                                      (None, T.int64), (None, T.decimal128(10, 2))])
def test_ipc_literals_of_every_type(value, dt):
    plan = PL.ProjectExec([(E.Literal(value, dt), "x")], PL.MemoryExec(S))
    txt = plan.explain()
    assert f":{dt}" in txt.splitlines()[0]
    if value is None:
        assert "NULL:" in txt
    elif dt.is_decimal:
        assert f"dec({-1 if value < 0 else 0}:{value % 2**64})" in txt
    elif not dt.is_float:
        assert f"{int(value)}:{dt}" in txt


def test_task_definition_kind():
    plan = PL.FilterExec([E.BinaryExpr(A, "Gt", B)], PL.MemoryExec(S))
    td = P.task_definition(plan.node(), 1, 2, 3)
    assert "FilterExec [(a@0 Gt b@1)]" in native.plan_explain(td, native.TASK_DEFINITION)


@pytest.mark.parametrize("build,code,msg", [
    (lambda: PL.FilterExec([], PL.MemoryExec(S)), native.ERR_INVALID_PLAN, "at least one predicate"),              # filter_exec.rs:58-60
    (lambda: PL.FilterExec([A], PL.MemoryExec(S)), native.ERR_INVALID_PLAN, "must return boolean"),                  # filter_exec.rs:61-66
    (lambda: PL.FilterExec([E.BinaryExpr(E.Column("zz"), "Lt", A)], PL.MemoryExec(S)), native.ERR_INVALID_PLAN, 'Unable to get field named "zz"'),
    (lambda: PL.FilterExec([E.BinaryExpr(A, "Lt", I)], PL.MemoryExec(S)), native.ERR_INVALID_PLAN, "equal types"),
])
def test_validation_errors(build, code, msg):
    with pytest.raises(native.NativeError) as ei:
        build()
    assert ei.value.code == code and msg in str(ei.value)


def test_unsupported_is_reported_not_faked():
    n = P.PhysicalPlanNode()
    n.filter.input.CopyFrom(P.ffi_reader_node(S))
    e = n.filter.expr.add()
    e.binary_expr.l.CopyFrom(P.expr_msg(A)); e.binary_expr.r.CopyFrom(P.expr_msg(B)); e.binary_expr.op = "RegexMatch"
    with pytest.raises(native.NativeError) as ei:
        native.plan_explain(n.SerializeToString())
    assert ei.value.code == native.ERR_UNSUPPORTED
    e.binary_expr.op = "Frobnicate"
    with pytest.raises(native.NativeError) as ei:
        native.plan_explain(n.SerializeToString())
    assert ei.value.code == native.ERR_INVALID_PLAN and "Unsupported binary operator" in str(ei.value)   # auron-serde lib.rs:97-100
    with pytest.raises(native.NativeError) as ei:
        native.plan_explain(b"\x42\x05\x0a")          # FilterExecNode with a truncated length-delimited body
    assert ei.value.code == native.ERR_INVALID_PLAN


def test_explain_shuffle_writer_and_partitionings():
    """ShuffleWriterExecNode + PhysicalRepartition decode (auron.proto:524-529,629-655; from_proto.rs:263-277,1107-1187)"""
    ins = T.Schema([T.Field("k", T.int64, False), T.Field("v", T.float64, True)])
    leaf = PL.MemoryExec(ins)
    h = PL.ShuffleWriterExec(PL.FilterExec([E.BinaryExpr(E.Column("k"), "Gt", E.Literal(3, T.int64))], leaf), ("hash", [E.Column("k")], 200), "/x/shuffle_0_0.data", "/x/shuffle_0_0.index")
    text = h.explain()
    assert "ShuffleWriterExec partitioning=Hash([k@0], 200) data=/x/shuffle_0_0.data index=/x/shuffle_0_0.index" in text and "FilterExec" in text
    assert "partitioning=Single([], 1)" in PL.ShuffleWriterExec(leaf, ("single",), "a", "b").explain()
    assert "partitioning=RoundRobin([], 9)" in PL.ShuffleWriterExec(leaf, ("round_robin", 9), "a", "b").explain()
    assert h.schema() == leaf.schema()                       # shuffle_writer_exec.rs:76-78
    with pytest.raises(native.NativeError) as ei:            # hash expressions resolve against the input schema (from_proto.rs:1122-1126)
        PL.ShuffleWriterExec(leaf, ("hash", [E.Column("nope")], 4), "a", "b")
    assert ei.value.code == native.ERR_INVALID_PLAN
