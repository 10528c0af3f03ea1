"""The reference's hash-join goldens (datafusion-ext-plans/src/joins/test.rs:275-985), restated as data: each case is
(name, left batches, right batches, on [(left col, right col)], join type, expected rows).  A table is a list of
batches; a batch is {column name: values}; None = NULL.  Column types are Int32 except where `dtype` says otherwise.
The reference runs every case under five drivers; the four hash-join ones (BHJ / SHJ, left- or right-probed) map to
build_side = "right" / "left" here.  Comparison is on the sorted rows (`assert_batches_sorted_eq!`)."""
from oracle import join_oracle as J


def t(a, b, c):
    return [dict([a, b, c])]


N = None
CASES = [
    ("join_inner_one", t(("a1", [1, 2, 3]), ("b1", [4, 5, 5]), ("c1", [7, 8, 9])), t(("a2", [10, 20, 30]), ("b1", [4, 5, 6]), ("c2", [70, 80, 90])),
     [("b1", "b1")], J.INNER, [(1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (3, 5, 9, 20, 5, 80)]),
    ("join_inner_two", t(("a1", [1, 2, 2]), ("b2", [1, 2, 2]), ("c1", [7, 8, 9])), t(("a1", [1, 2, 3]), ("b2", [1, 2, 2]), ("c2", [70, 80, 90])),
     [("a1", "a1"), ("b2", "b2")], J.INNER, [(1, 1, 7, 1, 1, 70), (2, 2, 8, 2, 2, 80), (2, 2, 9, 2, 2, 80)]),
    ("join_inner_two_two", t(("a1", [1, 1, 2]), ("b2", [1, 1, 2]), ("c1", [7, 8, 9])), t(("a1", [1, 1, 3]), ("b2", [1, 1, 2]), ("c2", [70, 80, 90])),
     [("a1", "a1"), ("b2", "b2")], J.INNER, [(1, 1, 7, 1, 1, 70), (1, 1, 7, 1, 1, 80), (1, 1, 8, 1, 1, 70), (1, 1, 8, 1, 1, 80)]),
    ("join_inner_with_nulls", t(("a1", [1, 1, 2, 2]), ("b2", [N, 1, 2, 2]), ("c1", [1, N, 8, 9])), t(("a1", [1, 1, 2, 3]), ("b2", [N, 1, 2, 2]), ("c2", [10, 70, 80, 90])),
     [("a1", "a1"), ("b2", "b2")], J.INNER, [(1, 1, N, 1, 1, 70), (2, 2, 8, 2, 2, 80), (2, 2, 9, 2, 2, 80)]),
    ("join_left_one", t(("a1", [1, 2, 3]), ("b1", [4, 5, 7]), ("c1", [7, 8, 9])), t(("a2", [10, 20, 30]), ("b1", [4, 5, 6]), ("c2", [70, 80, 90])),
     [("b1", "b1")], J.LEFT, [(1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (3, 7, 9, N, N, N)]),
    ("join_right_one", t(("a1", [1, 2, 3]), ("b1", [4, 5, 7]), ("c1", [7, 8, 9])), t(("a2", [10, 20, 30]), ("b1", [4, 5, 6]), ("c2", [70, 80, 90])),
     [("b1", "b1")], J.RIGHT, [(1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (N, N, N, 30, 6, 90)]),
    ("join_full_one", t(("a1", [1, 2, 2, 3]), ("b1", [4, 5, 5, 7]), ("c1", [7, 8, 80, 9])), t(("a2", [10, 20, 20, 30]), ("b2", [4, 5, 5, 6]), ("c2", [70, 80, 800, 90])),
     [("b1", "b2")], J.FULL, [(N, N, N, 30, 6, 90), (1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (2, 5, 8, 20, 5, 800), (2, 5, 80, 20, 5, 80), (2, 5, 80, 20, 5, 800), (3, 7, 9, N, N, N)]),
    ("join_anti", t(("a1", [1, 2, 2, 3, 5]), ("b1", [4, 5, 5, 7, 7]), ("c1", [7, 8, 8, 9, 11])), t(("a2", [10, 20, 30]), ("b1", [4, 5, 6]), ("c2", [70, 80, 90])),
     [("b1", "b1")], J.LEFT_ANTI, [(3, 7, 9), (5, 7, 11)]),
    ("join_semi", t(("a1", [1, 2, 2, 3]), ("b1", [4, 5, 5, 7]), ("c1", [7, 8, 8, 9])), t(("a2", [10, 20, 30]), ("b1", [4, 5, 6]), ("c2", [70, 80, 90])),
     [("b1", "b1")], J.LEFT_SEMI, [(1, 4, 7), (2, 5, 8), (2, 5, 8)]),
    ("join_with_duplicated_column_names", t(("a", [1, 2, 3]), ("b", [4, 5, 7]), ("c", [7, 8, 9])), t(("a", [10, 20, 30]), ("b", [1, 2, 7]), ("c", [70, 80, 90])),
     [("a", "b")], J.INNER, [(1, 4, 7, 10, 1, 70), (2, 5, 8, 20, 2, 80)]),
    ("join_date32", t(("a1", [1, 2, 3]), ("b1", [19107, 19108, 19108]), ("c1", [7, 8, 9])), t(("a2", [10, 20, 30]), ("b1", [19107, 19108, 19109]), ("c2", [70, 80, 90])),
     [("b1", "b1")], J.INNER, [(1, 19107, 7, 10, 19107, 70), (2, 19108, 8, 20, 19108, 80), (3, 19108, 9, 20, 19108, 80)], "date32"),
    # Date64 (milliseconds as i64) is outside this repo's type subset: the same vectors run as timestamp[us] — an 8-byte integer key either way
    ("join_date64", t(("a1", [1, 2, 3]), ("b1", [1650703441000, 1650903441000, 1650903441000]), ("c1", [7, 8, 9])),
     t(("a2", [10, 20, 30]), ("b1", [1650703441000, 1650503441000, 1650903441000]), ("c2", [70, 80, 90])),
     [("b1", "b1")], J.INNER, [(1, 1650703441000, 7, 10, 1650703441000, 70), (2, 1650903441000, 8, 30, 1650903441000, 90), (3, 1650903441000, 9, 30, 1650903441000, 90)], "timestamp"),
    ("join_left_sort_order", t(("a1", [0, 1, 2, 3, 4, 5]), ("b1", [3, 4, 5, 6, 6, 7]), ("c1", [4, 5, 6, 7, 8, 9])), t(("a2", [0, 10, 20, 30, 40]), ("b2", [2, 4, 6, 6, 8]), ("c2", [50, 60, 70, 80, 90])),
     [("b1", "b2")], J.LEFT, [(0, 3, 4, N, N, N), (1, 4, 5, 10, 4, 60), (2, 5, 6, N, N, N), (3, 6, 7, 20, 6, 70), (3, 6, 7, 30, 6, 80), (4, 6, 8, 20, 6, 70), (4, 6, 8, 30, 6, 80), (5, 7, 9, N, N, N)]),
    ("join_right_sort_order", t(("a1", [0, 1, 2, 3]), ("b1", [3, 4, 5, 7]), ("c1", [6, 7, 8, 9])), t(("a2", [0, 10, 20, 30]), ("b2", [2, 4, 5, 6]), ("c2", [60, 70, 80, 90])),
     [("b1", "b2")], J.RIGHT, [(N, N, N, 0, 2, 60), (1, 4, 7, 10, 4, 70), (2, 5, 8, 20, 5, 80), (N, N, N, 30, 6, 90)]),
]

_L2 = [dict([("a1", [0, 1, 2]), ("b1", [3, 4, 5]), ("c1", [4, 5, 6])]), dict([("a1", [3, 4, 5, 6]), ("b1", [6, 6, 7, 9]), ("c1", [7, 8, 9, 9])])]
_R2 = [dict([("a2", [0, 10, 20]), ("b2", [2, 4, 6]), ("c2", [50, 60, 70])]), dict([("a2", [30, 40]), ("b2", [6, 8]), ("c2", [80, 90])])]
CASES += [
    ("join_left_multiple_batches", _L2, _R2, [("b1", "b2")], J.LEFT,
     [(0, 3, 4, N, N, N), (1, 4, 5, 10, 4, 60), (2, 5, 6, N, N, N), (3, 6, 7, 20, 6, 70), (3, 6, 7, 30, 6, 80), (4, 6, 8, 20, 6, 70), (4, 6, 8, 30, 6, 80), (5, 7, 9, N, N, N), (6, 9, 9, N, N, N)]),
    ("join_right_multiple_batches",
     [dict([("a1", [0, 10, 20]), ("b1", [2, 4, 6]), ("c1", [50, 60, 70])]), dict([("a1", [30, 40]), ("b1", [6, 8]), ("c1", [80, 90])])],
     [dict([("a2", [0, 1, 2]), ("b2", [3, 4, 5]), ("c2", [4, 5, 6])]), dict([("a2", [3, 4, 5, 6]), ("b2", [6, 6, 7, 9]), ("c2", [7, 8, 9, 9])])],
     [("b1", "b2")], J.RIGHT,
     [(N, N, N, 0, 3, 4), (10, 4, 60, 1, 4, 5), (N, N, N, 2, 5, 6), (20, 6, 70, 3, 6, 7), (30, 6, 80, 3, 6, 7), (20, 6, 70, 4, 6, 8), (30, 6, 80, 4, 6, 8), (N, N, N, 5, 7, 9), (N, N, N, 6, 9, 9)]),
    ("join_full_multiple_batches", _L2, _R2, [("b1", "b2")], J.FULL,
     [(N, N, N, 0, 2, 50), (N, N, N, 40, 8, 90), (0, 3, 4, N, N, N), (1, 4, 5, 10, 4, 60), (2, 5, 6, N, N, N), (3, 6, 7, 20, 6, 70), (3, 6, 7, 30, 6, 80), (4, 6, 8, 20, 6, 70), (4, 6, 8, 30, 6, 80),
      (5, 7, 9, N, N, N), (6, 9, 9, N, N, N)]),
    ("join_existence_multiple_batches", _L2, _R2, [("b1", "b2")], J.EXISTENCE,
     [(0, 3, 4, False), (1, 4, 5, True), (2, 5, 6, False), (3, 6, 7, True), (4, 6, 8, True), (5, 7, 9, False), (6, 9, 9, False)]),
]


def arrow_batches(table, dtype="int32"):
    import pyarrow as pa
    nullable = any(v is None for b in table for col in b.values() for v in col)       # build_table_i32_nullable declares nullable fields (test.rs:147-160)
    base = {"int32": pa.int32(), "date32": pa.int32(), "timestamp": pa.int64()}[dtype]
    final = {"int32": pa.int32(), "date32": pa.date32(), "timestamp": pa.timestamp("us")}[dtype]
    names = list(table[0])
    schema = pa.schema([pa.field(n, final, nullable) for n in names])
    return [pa.RecordBatch.from_arrays([pa.array(b[n], base).cast(final) for n in names], schema=schema) for b in table]
