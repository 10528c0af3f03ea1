"""ShuffleWriterExec on the GPU through the C ABI vs the oracle (oracle/shuffle_oracle.py):
the .data / .index files are read back the way the reduce side does (IpcCompressionReader + read_batch) and every
partition must hold exactly the rows pmod(murmur3(keys, 42), n) sends there — as a multiset: the row order inside a
partition is not a contract (the reference's radix sort is unstable, rdx_sort.rs:55-73)."""
import decimal
import os
import struct

import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import exprs as E, native, plans as PL, types as T
from oracle import blaze_oracle as O
from oracle import shuffle_oracle as S
from helpers import *

pytestmark = pytest.mark.gpu


def _table(n, seed, null_frac):
    rng = np.random.default_rng(seed)
    k = rng.integers(-50, 5000, n, dtype=np.int64)
    k32 = rng.integers(-3, 40, n).astype(np.int32)
    cols = {
        "k": with_nulls(rng, k, null_frac / 4),
        "k32": with_nulls(rng, k32, null_frac, pa.int32()),
        "i8": with_nulls(rng, rng.integers(-128, 128, n).astype(np.int8), null_frac, pa.int8()),
        "i16": pa.array(rng.integers(-2**15, 2**15, n).astype(np.int16)),
        "f32": with_nulls(rng, rng.normal(size=n).astype(np.float32), null_frac, pa.float32()),
        "f64": with_nulls(rng, rng.normal(0, 1e9, n), null_frac),
        "d": pa.array(rng.integers(0, 20000, n).astype(np.int32), pa.int32()).cast(pa.date32()),
        "ts": with_nulls(rng, rng.integers(0, 2**50, n, dtype=np.int64), null_frac).cast(pa.timestamp("us")),
        "dec": pa.array([decimal.Decimal(int(x)).scaleb(-2) for x in rng.integers(-10**15, 10**15, n)], pa.decimal128(20, 2),
                        mask=(rng.random(n) < null_frac) if null_frac else None),
        "b": with_nulls(rng, rng.random(n) < 0.3, null_frac, pa.bool_()),
    }
    names = list(cols)
    fields = [pa.field(c, cols[c].type, c not in ("i16", "d")) for c in names]
    return pa.RecordBatch.from_arrays([cols[c] for c in names], schema=pa.schema(fields))


def _check_files(tmp_path, plan, batches, part, keys, P, batch_size):
    data = open(plan.output_data_file, "rb").read()
    index = open(plan.output_index_file, "rb").read()
    assert len(index) == 8 * (P + 1)                                             # sort_repartitioner.rs:181-185
    offs = struct.unpack("<%dq" % (P + 1), index)
    assert offs[0] == 0 and offs[-1] == len(data) and all(a <= b for a, b in zip(offs, offs[1:]))
    schema = T.from_arrow_schema(batches[0].schema)
    parts = S.read_shuffle_file(data, index, schema)
    whole = O.concat_batches(schema, oracle_batches(batches))
    pid = S.evaluate_partition_ids(part, whole) if P > 1 else np.zeros(whole.num_rows, np.uint32)
    total = 0
    for q in range(P):
        exp = whole.take(np.nonzero(pid == q)[0])
        assert O.rows_multiset(parts[q]) == O.rows_multiset([exp]), f"partition {q}"
        if exp.num_rows == 0:
            assert offs[q] == offs[q + 1]                                        # empty partitions take no bytes (buffered_data.rs:141-153)
        for b in parts[q]:
            assert 0 < b.num_rows <= batch_size
        total += sum(b.num_rows for b in parts[q])
    assert total == whole.num_rows
    # the chunks (bytes before compression) are what the files frame
    for q in range(P):
        raw = b"".join(ch["data"][ch["part_off"][q]: ch["part_off"][q + 1]] for ch in plan.last_chunks)
        assert raw == S.read_ipc_blocks(data[offs[q]: offs[q + 1]])
    assert sum(ch["rows"] for ch in plan.last_chunks) == whole.num_rows


@pytest.mark.parametrize("P,keys,null_frac,batch_size", [(7, ["k"], 0.0, 10000), (200, ["k", "k32"], 0.15, 10000), (1000, ["k32", "dec", "b"], 0.1, 64),
                                                         (3, ["f64", "i8", "ts"], 0.2, 100), (1, [], 0.1, 500)])
def test_hash_shuffle_files_match_the_oracle(tmp_path, P, keys, null_frac, batch_size):
    rb = _table(30_000, 21 + P, null_frac)
    batches = split_batches(rb, 7_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    part = ("hash", [E.Column(k) for k in keys], P) if P > 1 else ("single",)
    plan = PL.ShuffleWriterExec(leaf, part, str(tmp_path / "s.data"), str(tmp_path / "s.index"))
    assert PL.collect(plan, native.default_conf(staging_rows=0, batch_size=batch_size)) == []          # the stream is empty (shuffle_writer_exec.rs:109-165)
    m = plan.last_metrics
    assert m["fast_path_launches"] == len(batches) and m["gpu_kernel_launches"] >= 3 * len(batches)
    sp = S.Partitioning("hash", P, hash_cols=[rb.schema.names.index(k) for k in keys])
    _check_files(tmp_path, plan, batches, sp, keys, P, batch_size)


def test_staged_small_batches_and_a_fused_filter(tmp_path):
    """10,000-row host batches go through the pinned staging ring; a FilterExec below the writer runs as its own stage"""
    rb = _table(50_000, 5, 0.1)
    batches = split_batches(rb, 10_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    preds = [E.BinaryExpr(E.Column("k32"), "GtEq", E.Literal(5, T.int32))]
    plan = PL.ShuffleWriterExec(PL.FilterExec(preds, leaf), ("hash", [E.Column("k")], 16), str(tmp_path / "s.data"), str(tmp_path / "s.index"))
    PL.collect(plan, native.default_conf(staging_rows=1 << 15))
    ins = T.from_arrow_schema(rb.schema)
    kept = O.FilterExec(preds, ins).execute(oracle_batches(batches))
    kept_arrow = [O.batch_to_arrow(b) for b in kept]
    _check_files(tmp_path, plan, kept_arrow, S.Partitioning("hash", 16, hash_cols=[0]), ["k"], 16, 10000)


def test_partial_aggregate_feeds_the_writer(tmp_path):
    """AggExec(Partial, columnar states) -> ShuffleWriterExec: the q1 map side; the shuffled rows are the partial states"""
    rng = np.random.default_rng(8)
    n = 40_000
    rb = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 3000, n, dtype=np.int64)), pa.array(rng.integers(-10**6, 10**6, n, dtype=np.int64))], names=["k", "v"])
    batches = split_batches(rb, 10_000)
    leaf = PL.MemoryExec.from_arrow(batches, rb.schema)
    ins = leaf.schema()
    g = [E.GroupingExpr("k", E.Column("k"))]
    aggs = [E.AggExpr("s", E.PARTIAL, PL.create_agg(E.AGG_SUM, [E.Column("v")], ins, T.int64)), E.AggExpr("c", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Column("v")], ins, T.int64))]
    partial = PL.AggExec(PL.HashAgg, g, aggs, False, leaf, columnar_state=True)
    plan = PL.ShuffleWriterExec(partial, ("hash", [E.Column("k")], 32), str(tmp_path / "s.data"), str(tmp_path / "s.index"))
    PL.collect(plan, native.default_conf(partial_state_columnar=1))
    data, index = open(plan.output_data_file, "rb").read(), open(plan.output_index_file, "rb").read()
    parts = S.read_shuffle_file(data, index, partial.schema())
    k = rb.column(0).to_numpy(); v = rb.column(1).to_numpy()
    exp = {}
    for kk, vv in zip(k, v):
        s, c = exp.get(int(kk), (0, 0)); exp[int(kk)] = (s + int(vv), c + 1)
    seen = {}
    for q, bs in enumerate(parts):
        for b in bs:
            pid = S.evaluate_partition_ids(S.Partitioning("hash", 32, hash_cols=[0]), b)
            assert (pid == q).all()
            for r in range(b.num_rows):
                key = int(b.cols[0].values[r]); assert key not in seen
                seen[key] = (int(b.cols[1].values[r]), int(b.cols[2].values[r]))
    assert seen == exp


def test_chunks_can_stay_on_the_device(tmp_path):
    rb = _table(20_000, 3, 0.0)
    leaf = PL.MemoryExec.from_arrow([rb], rb.schema)
    plan = PL.ShuffleWriterExec(leaf, ("hash", [E.Column("k")], 50), str(tmp_path / "s.data"), str(tmp_path / "s.index"))
    PL.collect(plan, native.default_conf(staging_rows=0, shuffle_output_on_device=1))
    assert not os.path.exists(plan.output_data_file)
    (ch,) = plan.last_chunks
    assert ch["on_device"] and ch["data_ptr"] and ch["rows"] == 20_000 and sum(ch["part_rows"]) == 20_000
    whole = O.batch_from_arrow(rb)
    pid = S.evaluate_partition_ids(S.Partitioning("hash", 50, hash_cols=[0]), whole)
    assert ch["part_rows"] == [int((pid == q).sum()) for q in range(50)]


def test_partitionings_outside_the_gpu_path_are_refused(tmp_path):
    """round-robin needs the sort operator first (shuffle_writer_exec.rs:133-158): UNSUPPORTED at create, so the host keeps its CPU operator"""
    rb = _table(100, 1, 0.0)
    leaf = PL.MemoryExec.from_arrow([rb], rb.schema)
    plan = PL.ShuffleWriterExec(leaf, ("round_robin", 4), str(tmp_path / "a"), str(tmp_path / "b"))
    with pytest.raises(native.NativeError) as ei:
        PL.collect(plan)
    assert ei.value.code == native.ERR_UNSUPPORTED
    one = PL.ShuffleWriterExec(leaf, ("round_robin", 1), str(tmp_path / "a"), str(tmp_path / "b"))        # partition_count() == 1 -> SingleShuffleRepartitioner (:120-124)
    PL.collect(one)
    assert struct.unpack("<2q", open(one.output_index_file, "rb").read())[0] == 0


def test_empty_input_writes_empty_files(tmp_path):
    """buffered_data.rs:124-126 / single_repartitioner.rs:87-96: no rows -> empty .data, all-zero .index"""
    rb = _table(10, 1, 0.0).slice(0, 0)
    leaf = PL.MemoryExec.from_arrow([rb], rb.schema)
    plan = PL.ShuffleWriterExec(leaf, ("hash", [E.Column("k")], 5), str(tmp_path / "s.data"), str(tmp_path / "s.index"))
    PL.collect(plan)
    assert open(plan.output_data_file, "rb").read() == b"" and open(plan.output_index_file, "rb").read() == bytes(48)


def _np_murmur3_pid(k, parts):
    """vectorised pmod(murmur3(le_bytes(int64), 42), parts) (hash/mur.rs:19-87) — the pure-python oracle is too slow for millions of rows"""
    M = np.uint64(0xFFFFFFFF)
    k = k.astype(np.int64).view(np.uint64)
    def mul(a, b): return (a * np.uint64(b)) & M
    def rotl(x, r): return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & M
    def mix_k1(k1): return mul(rotl(mul(k1, 0xcc9e2d51), 15), 0x1b873593)
    def mix_h1(h1, k1): return (mul(rotl(h1 ^ k1, 13), 5) + np.uint64(0xe6546b64)) & M
    h = mix_h1(mix_h1(np.full(len(k), 42, np.uint64), mix_k1(k & M)), mix_k1(k >> np.uint64(32)))
    h ^= np.uint64(8); h ^= h >> np.uint64(16); h = mul(h, 0x85ebca6b); h ^= h >> np.uint64(13); h = mul(h, 0xc2b2ae35); h ^= h >> np.uint64(16)
    return np.mod(h.astype(np.int64).astype(np.int32).astype(np.int64), parts)


def test_many_tiles_per_cta_and_bulk_copy_staging(tmp_path):
    """2.6 M rows in one device-sized batch: every CTA walks several 4096-row tiles (the cp.async.bulk double buffering crosses
    tile boundaries), a partial last tile, 8 / 4 / 2-byte columns; every partition must hold exactly its rows"""
    rng = np.random.default_rng(12)
    n = 2_600_123
    k = rng.integers(-10**12, 10**12, n, dtype=np.int64)
    rb = pa.RecordBatch.from_arrays([pa.array(k), pa.array(np.arange(n, dtype=np.int64)), pa.array(rng.integers(-2**31, 2**31, n).astype(np.int32), pa.int32()),
                                     pa.array((np.arange(n) % 30000).astype(np.int16), pa.int16())],
                                    schema=pa.schema([pa.field("k", pa.int64(), False), pa.field("row", pa.int64(), False), pa.field("v", pa.int32(), False), pa.field("s", pa.int16(), False)]))
    P = 200
    plan = PL.ShuffleWriterExec(PL.MemoryExec.from_arrow([rb], rb.schema), ("hash", [E.Column("k")], P), str(tmp_path / "s.data"), str(tmp_path / "s.index"))
    PL.collect(plan, native.default_conf(staging_rows=0))
    data, index = open(plan.output_data_file, "rb").read(), open(plan.output_index_file, "rb").read()
    parts = S.read_shuffle_file(data, index, T.from_arrow_schema(rb.schema))
    pid = _np_murmur3_pid(k, P)
    assert [int(O.murmur3_long(int(x), 42)) % P for x in k[:50]] == list(pid[:50])          # the vectorised hash agrees with the oracle's
    v, s = rb.column(2).to_numpy(), rb.column(3).to_numpy()
    seen = np.zeros(n, bool)
    for q in range(P):
        rows = np.concatenate([b.cols[1].values for b in parts[q]]) if parts[q] else np.zeros(0, np.int64)
        assert (pid[rows] == q).all() and not seen[rows].any()
        seen[rows] = True
        assert np.array_equal(np.concatenate([b.cols[0].values for b in parts[q]]), k[rows])
        assert np.array_equal(np.concatenate([b.cols[2].values for b in parts[q]]), v[rows]) and np.array_equal(np.concatenate([b.cols[3].values for b in parts[q]]), s[rows])
    assert seen.all()
