"""Committed golden fixtures (tests/golden/*.arrow, *.npz; generator: tests/golden/make_golden.py).

CPU: the pinned oracle still reproduces every fixture (drift guard).
GPU: the CUDA path, through the C ABI, matches the committed files — the oracle is not run, only its row-comparison
helpers are used."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pyarrow as pa
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
import cases                      # noqa: E402
import make_golden                # noqa: E402
from blaze_b200 import plans as PL, native     # noqa: E402
from oracle import blaze_oracle as O           # noqa: E402
from helpers import assert_multiset_equal, assert_same_rows_ordered, split_batches   # noqa: E402

CASES = {c.name: c for c in cases.all_cases()}
MANIFEST = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))


def read_ipc(name):
    with pa.OSFile(os.path.join(GOLDEN, name), "rb") as f:
        t = pa.ipc.open_file(f).read_all()
    return t.combine_chunks().to_batches()[0] if t.num_rows else pa.RecordBatch.from_pylist([], schema=t.schema)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_committed_fixture(name):
    c = CASES[name]
    assert read_ipc(name + ".in.arrow").equals(c.rb), "seeded input changed: regenerate with tests/golden/make_golden.py"
    exp = read_ipc(name + ".out.arrow")
    assert exp.num_rows == MANIFEST[name]["rows_out"]
    now = make_golden.expected_table(c)
    assert now.schema.equals(exp.schema)
    if c.ordered:
        assert now.equals(exp)
    else:
        assert O.rows_multiset([O.batch_from_arrow(now)]) == O.rows_multiset([O.batch_from_arrow(exp)])


def test_oracle_reproduces_committed_murmur3_vectors():
    z = np.load(os.path.join(GOLDEN, "murmur3_partition.npz"))
    a, b, valid_b, nparts = cases.murmur3_case()
    assert np.array_equal(z["a"], a) and np.array_equal(z["b"], b) and np.array_equal(z["valid_b"], valid_b)
    h = O.create_murmur3_hashes([O.Col(O.T.int64, a, np.ones(len(a), bool)), O.Col(O.T.int32, b, valid_b)], len(a), 42)
    assert np.array_equal(z["hashes"], h)
    for n in nparts:
        assert np.array_equal(z["p%d" % n], O.partition_ids(h, n))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_committed_fixture(name):
    c = CASES[name]
    rb = read_ipc(name + ".in.arrow")
    leaf = PL.MemoryExec.from_arrow(split_batches(rb, c.batch_rows), rb.schema)
    plan = c.build(leaf)
    got = PL.collect(plan)
    exp = [O.batch_from_arrow(read_ipc(name + ".out.arrow"))]
    if c.ordered:
        assert_same_rows_ordered(got, exp, plan.schema())
    else:
        assert_multiset_equal(got, exp, c.float_cols)


@pytest.mark.gpu
def test_gpu_murmur3_partition_matches_committed_vectors():
    import torch
    z = np.load(os.path.join(GOLDEN, "murmur3_partition.npz"))
    a, b, valid_b = z["a"], z["b"], z["valid_b"]
    n = len(a)
    kids = [native.ArrowSchema() for _ in range(2)]
    for k, f in zip(kids, ["l", "i"]):
        k.format = f.encode(); k.name = b"c"
    arr = (C.POINTER(native.ArrowSchema) * 2)(*[C.pointer(k) for k in kids])
    top = native.ArrowSchema(); top.format = b"+s"; top.name = b""; top.n_children = 2
    top.children = C.cast(arr, C.POINTER(C.POINTER(native.ArrowSchema)))
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    tv = torch.from_numpy(np.packbits(valid_b, bitorder="little")).cuda()
    for nparts in (2, 7, 200):
        out = torch.empty(n, dtype=torch.int32, device="cuda")
        db = native.DeviceBatch([(ta.data_ptr(), 0, n), (tb.data_ptr(), tv.data_ptr(), n)], n, 0, keepalive=(ta, tb, tv))
        native.check(native.lib.b200q_murmur3_partition(C.addressof(top), C.addressof(db.dev), nparts, out.data_ptr(), None))
        torch.cuda.synchronize()
        native.DeviceBatch._live.pop(db._id, None)
        assert np.array_equal(out.cpu().numpy(), z["p%d" % nparts].astype(np.int32))
