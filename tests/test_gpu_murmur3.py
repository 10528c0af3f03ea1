"""b200q_murmur3_partition (device) vs the oracle's Spark-compatible murmur3 + pmod (pinned by the
reference's known-answer vectors in tests/test_oracle_golden.py)."""
import ctypes as C

import numpy as np
import pytest

from blaze_b200 import native, types as T
from oracle import blaze_oracle as O

pytestmark = pytest.mark.gpu


def _schema(fmts):
    kids = [native.ArrowSchema() for _ in fmts]
    for k, f in zip(kids, fmts):
        k.format = f.encode(); k.name = b"c"
    arr = (C.POINTER(native.ArrowSchema) * len(kids))(*[C.pointer(k) for k in kids])
    top = native.ArrowSchema(); top.format = b"+s"; top.name = b""; top.n_children = len(kids)
    top.children = C.cast(arr, C.POINTER(C.POINTER(native.ArrowSchema)))
    top._keep = (kids, arr)
    return top


@pytest.mark.parametrize("nparts", [2, 8, 200])
def test_partition_ids_match_spark_murmur3(nparts):
    import torch
    n = 100_003
    rng = np.random.default_rng(9)
    a = rng.integers(-2**62, 2**62, n, dtype=np.int64); b = rng.integers(-2**31, 2**31, n, dtype=np.int64).astype(np.int32)
    a[:5] = [1, 0, -1, 2**63 - 1, -2**63]
    valid_b = rng.random(n) >= 0.1
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    tv = torch.from_numpy(np.packbits(valid_b, bitorder="little")).cuda()
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    db = native.DeviceBatch([(ta.data_ptr(), 0, n), (tb.data_ptr(), tv.data_ptr(), n)], n, 0, keepalive=(ta, tb, tv))
    ks = _schema(["l", "i"])
    native.check(native.lib.b200q_murmur3_partition(C.addressof(ks), C.addressof(db.dev), nparts, out.data_ptr(), None))
    torch.cuda.synchronize()
    native.DeviceBatch._live.pop(db._id, None)
    cols = [O.Col(T.int64, a, np.ones(n, bool)), O.Col(T.int32, b, valid_b)]
    exp = O.partition_ids(O.create_murmur3_hashes(cols, n, 42), nparts)
    assert np.array_equal(out.cpu().numpy().astype(np.uint32), exp)
