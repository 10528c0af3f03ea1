"""Host side of the parquet scan without a GPU: the Thrift footer parser against pyarrow's metadata of the same file, the raw
Snappy decoder against pyarrow's encoder, plan decoding of ParquetScanExecNode (auron.proto:368-419)."""
import struct

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from blaze_b200 import exprs as E, native, plans as PL, types as T


def _footer(path):
    data = open(path, "rb").read()
    assert data[:4] == b"PAR1" and data[-4:] == b"PAR1"
    (n,) = struct.unpack("<I", data[-8:-4])
    return data[-8 - n:-8]


def test_footer_parser_agrees_with_libparquet(tmp_path):
    rng = np.random.default_rng(1)
    n = 20_000
    t = pa.table({"a": pa.array(rng.integers(0, 100, n, dtype=np.int64)), "b": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.2),
                  "d": pa.array(rng.integers(0, 1000, n).astype(np.int32), pa.int32()).cast(pa.date32()), "s": pa.array(["x%d" % i for i in range(n)]),
                  "ts": pa.array(rng.integers(0, 2**40, n, dtype=np.int64)).cast(pa.timestamp("us")), "i8": pa.array(rng.integers(-5, 5, n).astype(np.int8), pa.int8())})
    path = str(tmp_path / "t.parquet")
    pq.write_table(t, path, compression="snappy", row_group_size=6_000)
    text = native.parquet_explain(_footer(path))
    md = pq.ParquetFile(path).metadata
    assert f"rows={md.num_rows} flat=true" in text
    for name, arrow in (("a", "int64"), ("b", "float64"), ("d", "date32"), ("ts", "timestamp[us]"), ("i8", "int8"), ("s", "unsupported")):
        assert any(l.startswith(f"column {name} ") and l.endswith(f"arrow={arrow}") for l in text.splitlines()), (name, text)
    lines = [l.strip() for l in text.splitlines()]
    for g in range(md.num_row_groups):
        rg = md.row_group(g)
        assert f"row_group {g} rows={rg.num_rows}" in lines
        for c in range(rg.num_columns):
            col = rg.column(c)
            start = col.dictionary_page_offset if col.dictionary_page_offset and col.dictionary_page_offset < col.data_page_offset else col.data_page_offset
            want = f"chunk {c} codec=1 values={col.num_values} start={start} bytes={col.total_compressed_size} nulls={col.statistics.null_count}"
            assert any(l.startswith(want) for l in lines), (want, text)


@pytest.mark.parametrize("kind", ["empty", "text", "random", "runs", "large"])
def test_snappy_decoder_reads_what_a_conforming_encoder_writes(kind):
    rng = np.random.default_rng(4)
    data = {"empty": b"", "text": b"the quick brown fox jumps over the lazy dog " * 3000, "random": rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes(),
            "runs": bytes(100_000) + b"\x01" * 70_000 + bytes(range(256)) * 300, "large": (np.arange(400_000) // 9).astype(np.int64).tobytes()}[kind]
    comp = pa.compress(data, codec="snappy", asbytes=True)
    assert native.snappy_uncompress(comp, len(data)) == data


def test_snappy_decoder_fuzz_against_the_conforming_codec():
    # the decoder's fast loop moves 16 / 8-byte blocks with slop: many shapes (short periods = overlapping copies, long literals, column-like
    # integers, sizes around the fast-loop margins) against pyarrow's codec; truncated and bit-flipped streams must fail or decode, never crash
    rng = np.random.default_rng(12)
    shapes = []
    for n in [0, 1, 5, 15, 16, 17, 20, 21, 22, 63, 64, 65, 79, 80, 81, 100, 1000, 4097, 70_001]:
        shapes.append(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        shapes.append(bytes(rng.integers(0, 4, n, dtype=np.uint8)))
        for period in (1, 2, 3, 7, 8, 9, 15, 16, 17):
            shapes.append((bytes(rng.integers(0, 256, period, dtype=np.uint8)) * (n // period + 1))[:n])
    for bits, dt in ((21, np.int64), (11, np.int32), (40, np.int64), (7, np.int16)):
        shapes.append(rng.integers(0, 1 << bits, 50_000).astype(dt).tobytes())
        shapes.append(np.sort(rng.integers(0, 1 << bits, 50_000)).astype(dt).tobytes())
    for data in shapes:
        comp = pa.compress(data, codec="snappy", asbytes=True)
        assert native.snappy_uncompress(comp, len(data)) == data
    data = rng.integers(0, 1 << 21, 4000).astype(np.int64).tobytes()
    comp = bytearray(pa.compress(data, codec="snappy", asbytes=True))
    for cut in range(1, len(comp), 37):                                          # truncations: an error, never a crash
        with pytest.raises(native.NativeError):
            native.snappy_uncompress(bytes(comp[:cut]), len(data))
    for _ in range(300):                                                         # bit flips: either an error or SOME output of the declared size
        c = bytearray(comp); i = int(rng.integers(1, len(c))); c[i] ^= 1 << int(rng.integers(0, 8))
        try:
            native.snappy_uncompress(bytes(c), len(data) + 64)
        except native.NativeError:
            pass


def test_corrupt_inputs_fail_cleanly():
    with pytest.raises(native.NativeError):
        native.snappy_uncompress(b"\xff\xff\xff\xff\x0f\x00", 16)               # claims 4 GiB, then a literal that overruns
    with pytest.raises(native.NativeError):
        native.parquet_explain(b"\x15\x00\x19")                                    # truncated thrift


def test_parquet_scan_plan_decoding(tmp_path):
    ins = T.Schema([T.Field("k", T.int64, False), T.Field("v", T.float64, True), T.Field("d", T.date32, True)])
    pred = E.BinaryExpr(E.Column("k"), "Gt", E.Literal(10, T.int64))
    scan = PL.ParquetScanExec(ins, [("/data/part-0.parquet", 1000, (0, 500)), ("file:///data/part-1.parquet", 2000, None)], projection=[2, 0], pruning_predicates=[pred], limit=99)
    text = scan.explain()
    assert "ParquetScan files=[/data/part-0.parquet[0,500), file:///data/part-1.parquet] pruning=[(k@0 Gt 10:int64)] limit=99 schema=[d:date32?, k:int64]" in text
    agg = PL.AggExec(PL.HashAgg, [E.GroupingExpr("k", E.Column("k"))], [E.AggExpr("c", E.PARTIAL, PL.create_agg(E.AGG_COUNT, [E.Column("d")], scan.schema(), T.int64))], False,
                     PL.FilterExec([pred], scan))
    assert "AggExec" in agg.explain() and "ParquetScan" in agg.explain()
