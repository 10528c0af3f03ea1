"""The ShuffleWriterExec restatement (oracle/shuffle_oracle.py) against the reference's own goldens:
buffered_data.rs:394-540 (test_round_robin, test_range_partition, test_range_partition_2 — they pin partition ids AND the
row order the unstable American-flag sort leaves), rdx_sort.rs:81-114 (sortedness fuzz), batch_serde.rs:662-713 and
ipc_compression.rs:325-351 (round trips)."""
import struct

import numpy as np
import pyarrow as pa

from blaze_b200 import types as T
from oracle import blaze_oracle as O
from oracle import shuffle_oracle as S


def _table_i32():
    a = [19, 18, 17, 16, 15, 14, 13, 12, 11, 10]
    b = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
    c = [5, 6, 7, 8, 9, 0, 1, 2, 3, 4]
    rb = pa.RecordBatch.from_arrays([pa.array(x, pa.int32()) for x in (a, b, c)],
                                    schema=pa.schema([pa.field(n, pa.int32(), False) for n in "abc"]))
    return O.batch_from_arrow(rb)


def _rows(batch):
    return [tuple(int(c.values[r]) for c in batch.cols) for r in range(batch.num_rows)]


def test_reference_golden_round_robin():
    """buffered_data.rs:394-424"""
    _, sb = S.sort_batches_by_partition_id([_table_i32()], S.Partitioning("round_robin", 4), 3, 0)
    assert _rows(sb) == [(18, 1, 6), (14, 5, 0), (10, 9, 4), (17, 2, 7), (13, 6, 1), (12, 7, 2), (16, 3, 8), (19, 0, 5), (15, 4, 9), (11, 8, 3)]


def test_reference_golden_range_partition():
    """buffered_data.rs:426-478: one ascending Int32 key, bounds 11, 14, 17"""
    p = S.Partitioning("range", 4, sort_keys=[S.SortKey(0)], bounds=[(11,), (14,), (17,)])
    offs, sb = S.sort_batches_by_partition_id([_table_i32()], p, 0, 0)
    assert _rows(sb) == [(11, 8, 3), (10, 9, 4), (14, 5, 0), (13, 6, 1), (12, 7, 2), (17, 2, 7), (16, 3, 8), (15, 4, 9), (19, 0, 5), (18, 1, 6)]
    assert offs == [0, 2, 5, 8, 10]


def test_reference_golden_range_partition_2():
    """buffered_data.rs:480-540: two ascending keys, bounds (11,1), (14,3), (17,5)"""
    p = S.Partitioning("range", 4, sort_keys=[S.SortKey(0), S.SortKey(1)], bounds=[(11, 1), (14, 3), (17, 5)])
    _, sb = S.sort_batches_by_partition_id([_table_i32()], p, 0, 0)
    assert _rows(sb) == [(10, 9, 4), (13, 6, 1), (12, 7, 2), (11, 8, 3), (17, 2, 7), (16, 3, 8), (15, 4, 9), (14, 5, 0), (19, 0, 5), (18, 1, 6)]


def test_radix_sort_sorts_like_the_reference_fuzz():
    """rdx_sort.rs:81-114: same multiset, sorted by key; counts = bucket sizes"""
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 17, 1000):
        arr = [int(x) for x in rng.integers(0, 64, n)]
        got = list(arr)
        counts = S.radix_sort_by_key(got, 64, lambda k: k)
        assert got == sorted(arr)
        assert counts == [arr.count(k) for k in range(64)]


def test_binary_search_partitioning_above_128_bounds():
    """shuffle/mod.rs:234-275: > 128 bounds take the binary search; equal keys go to the bound's own partition"""
    bounds = [(10 * i,) for i in range(1, 200)]
    keys = [S.SortKey(0)]
    for v, exp in ((5, 0), (10, 0), (11, 1), (1990, 198), (1991, 199), (10**6, 199)):
        assert S.get_partition((v,), bounds, keys) == exp
        assert S.get_partition((v,), bounds[:100], keys) == min(exp, 100)           # linear scan form agrees


def _mixed_batch():
    """the arrays of batch_serde.rs:662-690, Binary instead of Utf8 (same wire format, batch_serde.rs:127-131)"""
    s = pa.array([b"20220101", "20220102你好🍹".encode(), "你好🍹20220103".encode(), None], pa.binary())
    u = pa.array([1000, 2000, 3000, None], pa.int64())
    bl = pa.array([True, False, None, None], pa.bool_())
    d = pa.array([None, 1, -2, 3], pa.int64()).cast(pa.decimal128(20, 0))
    f = pa.array([1.5, None, -0.0, float("inf")], pa.float64())
    i16 = pa.array([1, -2, None, 32767], pa.int16())
    i8 = pa.array([1, -2, None, 127], pa.int8())
    return pa.RecordBatch.from_arrays([s, u, bl, d, f, i16, i8], names=["str", "u64", "bool", "dec", "f", "i16", "i8"])


def test_batch_serde_round_trip_and_slice():
    """batch_serde.rs:662-713"""
    rb = _mixed_batch()
    for part in (rb, rb.slice(1, 2), rb.slice(0, 0)):
        b = O.batch_from_arrow(part)
        buf = S.write_batch(b.num_rows, b.cols)
        back, pos = S.read_batch(buf, 0, b.schema)
        assert pos == len(buf)
        assert O.rows_multiset([back]) == O.rows_multiset([b]) and back.num_rows == b.num_rows
        for c0, c1 in zip(b.cols, back.cols):
            assert list(c0.valid) == list(c1.valid)


def test_batch_serde_layout_of_a_primitive_column():
    """byte planes (transpose, batch_serde.rs:286-299), null-bit repacking (:205-223) and the varint header (io/mod.rs:60-69)"""
    vals = np.array([0x0102030405060708, -1, 0x1122334455667788], np.int64)
    c = O.Col(T.int64, vals, np.array([True, False, True]))
    buf = S.write_batch(3, [c])
    assert buf[0] == 3 and buf[1] == 1 and buf[2] == 0b101
    planes = buf[3:]
    assert planes[0:3] == bytes([0x08, 0x00, 0x88]) and planes[21:24] == bytes([0x01, 0x00, 0x11])       # plane 0 and plane 7; the NULL slot is stored as 0
    assert S.write_len(300) == bytes([128 + 44, 2])
    # 1-byte types are not transposed, booleans are bits
    assert S.write_array(O.Col(T.int8, np.array([1, 2, 3], np.int8), np.ones(3, bool))) == bytes([0, 1, 2, 3])
    assert S.write_array(O.Col(T.bool_, np.array([True, False, True]), np.ones(3, bool))) == bytes([0, 0b101])


def test_ipc_compression_round_trip():
    """ipc_compression.rs:325-351: two batches in one block"""
    w = S.IpcCompressionWriter()
    a = O.Col(T.binary, np.array([b"hello", b"world"], object), np.ones(2, bool))
    b = O.Col(T.binary, np.array([b"foo", b"bar"], object), np.ones(2, bool))
    w.write_batch(2, [a]); w.write_batch(2, [b]); w.finish_current_buf()
    data = bytes(w.out)
    (blen,) = struct.unpack_from("<I", data, 0)
    assert blen == len(data) - 4
    schema = T.Schema([T.Field("", T.binary, False)])
    got = S.read_partition(data, schema)
    assert [list(x.cols[0].values) for x in got] == [[b"hello", b"world"], [b"foo", b"bar"]]


def test_shuffle_write_hash_partitions_and_index():
    """sort_repartitioner.rs:151-185 + buffered_data.rs:123-158: every row lands in pmod(murmur3(keys, 42), n); the index
    holds n+1 little-endian i64 offsets; empty partitions take no bytes"""
    rng = np.random.default_rng(3)
    n = 5000
    rb = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 50, n), pa.int64()), pa.array(rng.integers(-9, 9, n), pa.int32()),
                                     pa.array(rng.normal(size=n), pa.float64(), mask=rng.random(n) < 0.1)], names=["k", "j", "x"])
    batches = [O.batch_from_arrow(rb.slice(i, 700)) for i in range(0, n, 700)]
    p = S.Partitioning("hash", 64, hash_cols=[0, 1])
    data, index = S.shuffle_write(batches, p, partition_id=7)
    assert len(index) == 8 * 65
    parts = S.read_shuffle_file(data, index, batches[0].schema)
    whole = O.concat_batches(batches[0].schema, batches)
    pid = S.evaluate_partition_ids(p, whole)
    for q in range(64):
        exp = whole.take(np.nonzero(pid == q)[0])
        assert O.rows_multiset(parts[q]) == O.rows_multiset([exp])
    offs = struct.unpack("<65q", index)
    assert offs[0] == 0 and offs[-1] == len(data) and all(a <= b for a, b in zip(offs, offs[1:]))
