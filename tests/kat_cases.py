"""Known-answer tests the REFERENCE's own unit tests hold for the cast / Spark-decimal helpers on the hot path
(SURVEY.md §8c), restated as (input column, projection, expected column).  Used twice: against the numpy oracle on
the CPU (tests/test_oracle_golden.py) and against the CUDA path through the C ABI (tests/test_gpu_reference_kats.py).

String-typed vectors of the same reference test modules (test_boolean_to_string, test_string_to_decimal,
test_decimal_to_string, test_string_to_bigint, test_string_to_date, TryCastExpr test_ok_2/3) are out of scope:
strings are not on the hot path."""
import pyarrow as pa

from blaze_b200 import exprs as E, types as T

I32_MAX, I32_MIN = 2**31 - 1, -2**31
X = E.Column("x")


def _dec(vals, p, s):
    return raw_decimal_array(vals, p, s)


def cases():
    """[(name, reference file:line, input pa.Array, projection expr, expected pa.Array)]"""
    out = []
    # datafusion-ext-commons/src/arrow/cast.rs:442-470 test_float_to_int: Rust `as` — truncation, saturation, NaN -> 0
    out.append(("cast f64->i32 (test_float_to_int)", "datafusion-ext-commons/src/arrow/cast.rs:442-470",
                pa.array([None, 123.456, 987.654, float(I32_MAX) + 10000.0, float(I32_MIN) - 10000.0, float("inf"), float("-inf"), float("nan")], pa.float64()),
                E.TryCast(X, T.int32),
                pa.array([None, 123, 987, I32_MAX, I32_MIN, I32_MAX, I32_MIN, 0], pa.int32())))
    # :472-492 test_int_to_float
    out.append(("cast i32->f64 (test_int_to_float)", "datafusion-ext-commons/src/arrow/cast.rs:472-492",
                pa.array([None, 123, 987, I32_MAX, I32_MIN], pa.int32()),
                E.TryCast(X, T.float64),
                pa.array([None, 123.0, 987.0, float(I32_MAX), float(I32_MIN)], pa.float64())))
    # :494-516 test_int_to_decimal: i32 -> Decimal128(38,18)
    out.append(("cast i32->decimal128(38,18) (test_int_to_decimal)", "datafusion-ext-commons/src/arrow/cast.rs:494-516",
                pa.array([None, 123, 987, I32_MAX, I32_MIN], pa.int32()),
                E.TryCast(X, T.decimal128(38, 18)),
                _dec([None, 123 * 10**18, 987 * 10**18, I32_MAX * 10**18, I32_MIN * 10**18], 38, 18)))
    # datafusion-ext-exprs/src/cast.rs:117-160 test_ok_1: TryCastExpr Float32 -> Int32
    out.append(("TryCastExpr f32->i32 (test_ok_1)", "datafusion-ext-exprs/src/cast.rs:117-160",
                pa.array([7.6, 9.0, 3.4, -0.0, -99.9, None], pa.float32()),
                E.TryCast(X, T.int32),
                pa.array([7, 9, 3, 0, -99, None], pa.int32())))
    # datafusion-ext-functions/src/spark_make_decimal.rs:73-101 test_decimal: i64 -> Decimal128(10,5), no range check
    md_vals = [12342132145623, 13245, 123213244568923, 1234567890, None]
    out.append(("MakeDecimal(10,5) (test_decimal)", "datafusion-ext-functions/src/spark_make_decimal.rs:73-101",
                pa.array(md_vals, pa.int64()),
                E.ScalarFunction("MakeDecimal", [X, E.Literal(10, T.int32), E.Literal(5, T.int32)], T.decimal128(10, 5)),
                "raw-decimal", (md_vals, 10, 5)))
    # datafusion-ext-functions/src/spark_unscaled_value.rs:54-77 test_unscaled_value_array
    uv_vals = [1234567890987654321, 9876543210, 135792468109, None, 67898]
    out.append(("UnscaledValue array (test_unscaled_value_array)", "datafusion-ext-functions/src/spark_unscaled_value.rs:54-77",
                ("raw-decimal", (uv_vals, 10, 8)),
                E.ScalarFunction("UnscaledValue", [X], T.int64),
                pa.array(uv_vals, pa.int64())))
    # :79-90 test_unscaled_value_scalar: Decimal128(123, 3, 2) -> 123 (the scalar broadcast over the batch)
    out.append(("UnscaledValue scalar (test_unscaled_value_scalar)", "datafusion-ext-functions/src/spark_unscaled_value.rs:79-90",
                pa.array([1], pa.int32()),
                E.ScalarFunction("UnscaledValue", [E.Literal(123, T.decimal128(3, 2))], T.int64),
                pa.array([123], pa.int64())))
    # datafusion-ext-functions/src/spark_check_overflow.rs:134-158: (20,8) -> (10,5), round half up, NULL on overflow
    out.append(("CheckOverflow (20,8)->(10,5) (test_check_overflow)", "datafusion-ext-functions/src/spark_check_overflow.rs:134-158",
                _dec([12342132145623, 13245, 123213244568923, 1234567890, None], 20, 8),
                E.ScalarFunction("CheckOverflow", [X, E.Literal(10, T.int32), E.Literal(5, T.int32)], T.decimal128(10, 5)),
                _dec([None, 13, None, 1234568, None], 10, 5)))
    return [normalise(c) for c in out]


def raw_decimal_array(vals, p, s):
    """Decimal128 array holding the given UNSCALED integers without any precision check (the reference's
    Decimal128Array::from(..).with_precision_and_scale only validates in debug paths; MakeDecimal does no range check)."""
    import numpy as np
    n = len(vals)
    data = b"".join((0 if v is None else v).to_bytes(16, "little", signed=True) for v in vals)
    valid = np.packbits(np.array([v is not None for v in vals], dtype=np.uint8), bitorder="little").tobytes()
    return pa.Array.from_buffers(pa.decimal128(p, s), n, [pa.py_buffer(valid), pa.py_buffer(data)], null_count=sum(v is None for v in vals))


def normalise(c):
    name, ref, inp, expr = c[0], c[1], c[2], c[3]
    exp = c[4:]
    if isinstance(inp, tuple) and inp[0] == "raw-decimal":
        inp = raw_decimal_array(*inp[1])
    if exp[0] == "raw-decimal":
        exp = raw_decimal_array(*exp[1])
    else:
        exp = exp[0]
    return name, ref, inp, expr, exp


def same_column(got: pa.Array, exp: pa.Array) -> bool:
    """validity equal and the valid values equal bit for bit (decimals: unscaled integers)"""
    import numpy as np
    if got.type != exp.type or len(got) != len(exp):
        return False
    gv, ev = np.array(got.is_valid()), np.array(exp.is_valid())
    if not np.array_equal(gv, ev):
        return False
    for i in range(len(got)):
        if not gv[i]:
            continue
        if pa.types.is_decimal(got.type):
            w = 16
            a = got.buffers()[1].to_pybytes()[(got.offset + i) * w:(got.offset + i + 1) * w]
            b = exp.buffers()[1].to_pybytes()[(exp.offset + i) * w:(exp.offset + i + 1) * w]
            if a != b:
                return False
        elif pa.types.is_floating(got.type):
            import struct
            f = "<d" if pa.types.is_float64(got.type) else "<f"
            if struct.pack(f, got[i].as_py()) != struct.pack(f, exp[i].as_py()):
                return False
        elif got[i].as_py() != exp[i].as_py():
            return False
    return True
