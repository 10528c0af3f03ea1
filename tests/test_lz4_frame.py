"""The library's LZ4 frame encoder (blaze_b200/csrc/lz4_frame.cc: the compression blocks of the shuffle files,
ipc_compression.rs:34-112) read back by an independent decoder (liblz4's LZ4F via pyarrow) — host only, no GPU."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import native


def _decode(frame: bytes) -> bytes:
    return pa.CompressedInputStream(pa.BufferReader(frame), "lz4").read()


@pytest.mark.parametrize("name", ["empty", "one byte", "twelve bytes", "thirteen bytes", "text", "random", "byte planes", "zeros above one block", "random above one block"])
def test_frames_decode_with_liblz4(name):
    rng = np.random.default_rng(7)
    data = {"empty": b"", "one byte": b"x", "twelve bytes": b"abcabcabcabc", "thirteen bytes": b"abcabcabcabca", "text": b"the quick brown fox " * 5000,
            "random": rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes(),
            "byte planes": np.frombuffer(rng.integers(-10**6, 10**6, 100_000, dtype=np.int64).tobytes(), np.uint8).reshape(-1, 8).T.tobytes(),
            "zeros above one block": bytes(9 << 20), "random above one block": rng.integers(0, 256, (4 << 20) + 12345, dtype=np.uint8).tobytes()}[name]
    frame = native.lz4_frame_compress(data)
    assert frame[:4] == bytes([0x04, 0x22, 0x4D, 0x18]) and frame[-4:] == bytes(4)          # magic, EndMark
    assert _decode(frame) == data
    if name in ("text", "zeros above one block"):
        assert len(frame) < len(data) // 20                                                  # it does compress
    if name == "byte planes":
        assert len(frame) < 0.8 * len(data)                                                  # the high-order planes (sign bytes) shrink, the low-order ones are noise
    if name.startswith("random"):
        assert len(frame) <= len(data) + 7 + 4 + 4 * (len(data) // (4 << 20) + 1)            # incompressible data is stored, never expanded


def test_every_match_length_and_offset_class():
    """token / extension-byte boundaries of the block format: literal runs and matches of 14..16, 269..271, 524.. bytes"""
    rng = np.random.default_rng(9)
    for lit in (0, 1, 14, 15, 16, 269, 270, 271, 600):
        for mlen in (4, 18, 19, 20, 273, 274, 275, 5000):
            pre = rng.integers(0, 256, lit + 64, dtype=np.uint8).tobytes()
            data = pre + pre[-min(len(pre), 40):] * (mlen // 40 + 1) + rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
            assert _decode(native.lz4_frame_compress(data)) == data


def test_too_small_output_buffer_reports_the_size():
    need = C.c_size_t(0)
    buf = C.create_string_buffer(4)
    st = native.lib.b200q_lz4_frame_compress(b"hello world, hello world", 24, buf, 4, C.byref(need))
    assert st == native.ERR_INVALID_ARG and need.value > 4


def test_shuffle_chunk_struct_layout():
    assert C.sizeof(native.ShuffleChunk) == 8 + 4 + 4 + 8 + 8 + 8
