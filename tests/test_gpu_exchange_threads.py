"""b200q_exchange_shuffle with the ranks of one job run as THREADS of this process — only on the emulated device
(tools/emu: the NCCL stand-in delivers sends between threads); on real GPUs one process drives one GPU and the same
entry point is covered by tests/test_gpu_multi.py (2 GPUs, torchrun).  Checks: every row lands on the rank
pmod(murmur3(keys, 42), world) names (oracle), nothing is lost or duplicated, values and validity travel with their keys."""
import threading

import numpy as np
import pyarrow as pa
import pytest

from blaze_b200 import native, types as T
from oracle import blaze_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif("emu" not in native.LIB_PATH, reason="ranks-as-threads needs the emulated device (tools/emu/run_gpu_suite.py)")]


def _np_col(dev_array, i, dtype):
    c = dev_array.array.children[i].contents
    n = c.length
    vals = np.ctypeslib.as_array((native.C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(c.buffers[1])).view(dtype).copy() if n else np.zeros(0, dtype)
    if c.buffers[0]:
        bits = np.ctypeslib.as_array((native.C.c_uint8 * ((n + 7) // 8)).from_address(c.buffers[0])) if n else np.zeros(0, np.uint8)
        valid = np.unpackbits(bits, bitorder="little")[:n].astype(bool)
    else:
        valid = np.ones(n, bool)
    return vals, valid


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("two_keys", [False, True])
def test_shuffle_between_thread_ranks(world, two_keys):
    uid = native.exchange_unique_id()
    results, errors = [None] * world, []
    schema = pa.schema([pa.field("k", pa.int64(), True)] + ([pa.field("k2", pa.int32(), False)] if two_keys else []) +
                       [pa.field("s", pa.int64(), True), pa.field("c", pa.int64(), False)])
    nk = 2 if two_keys else 1
    inputs = []
    for r in range(world):
        rng = np.random.default_rng(500 + r)
        n = [0, 1, 7000, 12345][r % 4] if world > 2 else 9000 + r
        k = rng.integers(-50, 400, n, dtype=np.int64); kvalid = rng.random(n) > 0.02
        k2 = rng.integers(0, 5, n).astype(np.int32)
        s = rng.integers(-10**9, 10**9, n, dtype=np.int64); svalid = rng.random(n) > 0.1
        c = rng.integers(0, 50, n, dtype=np.int64)
        inputs.append((k, kvalid, k2, s, svalid, c))

    def run(r):
        try:
            k, kvalid, k2, s, svalid, c = inputs[r]
            n = len(k)
            kb, sb = np.packbits(kvalid, bitorder="little"), np.packbits(svalid, bitorder="little")
            kb = np.concatenate([kb, np.zeros(8, np.uint8)]); sb = np.concatenate([sb, np.zeros(8, np.uint8)])
            cols = [(k.ctypes.data, kb.ctypes.data, n)] + ([(k2.ctypes.data, 0, n)] if two_keys else []) + [(s.ctypes.data, sb.ctypes.data, n), (c.ctypes.data, 0, n)]
            with native.Exchange(uid, r, world, 0) as ex:
                db = native.DeviceBatch(cols, n, 0, keepalive=(k, kb, k2, s, sb, c))
                out = ex.shuffle(schema, db.dev, nk)
                i = 0
                rk, rkv = _np_col(out, i, np.int64); i += 1
                rk2 = _np_col(out, i, np.int32)[0] if two_keys else None; i += two_keys
                rs, rsv = _np_col(out, i, np.int64); rc, _ = _np_col(out, i + 1, np.int64)
                native.release_device_array(out)
                assert ex.kernel_launches() > 0 or n == 0
            results[r] = (rk, rkv, rk2, rs, rsv, rc)
        except Exception as e:                                         # noqa: BLE001
            errors.append((r, repr(e)))

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors

    def rows(k, kv, k2, s, sv, c):
        return [(int(k[i]) if kv[i] else None, int(k2[i]) if k2 is not None else None, int(s[i]) if sv[i] else None, int(c[i])) for i in range(len(k))]

    sent = sorted(sum((rows(k, kv, k2 if two_keys else None, s, sv, c) for k, kv, k2, s, sv, c in inputs), []), key=repr)
    got = sorted(sum((rows(*res) for res in results), []), key=repr)
    assert got == sent                                                   # nothing lost, duplicated or altered
    for r, (rk, rkv, rk2, rs, rsv, rc) in enumerate(results):            # ownership = the reference's partitioning rule
        cols = [O.Col(T.int64, np.where(rkv, rk, 0), rkv)] + ([O.Col(T.int32, rk2, np.ones(len(rk2), bool))] if two_keys else [])
        pid = O.partition_ids(O.create_murmur3_hashes(cols, len(rk), 42), world)
        assert np.all(pid == r)
