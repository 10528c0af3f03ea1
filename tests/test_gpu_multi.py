"""2-GPU path (when the box has >= 2 GPUs): Partial per GPU -> murmur3(42) pmod ownership -> NCCL all_to_all of
the columnar partial states -> Final per GPU; the union of the ranks' results must equal the single-process
aggregate computed by the C restatement of the reference algorithm."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import ctypes as C
    import torch
    import torch.distributed as dist
    from blaze_b200 import exprs as E, native, plans as PL, types as T
    from blaze_b200.exchange import exchange_columns
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(7 + rank)
    k = rng.integers(0, 50_000, n, dtype=np.int64); v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    tk, tv = torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda()
    plans = bench.m1_plans()
    conf = native.default_conf(partial_state_columnar=1)
    with native.NativeOp(plans["partial_col"], conf, rank) as op:
        op.push_device(native.DeviceBatch([(tk.data_ptr(), 0, n), (tv.data_ptr(), 0, n)], n, rank, keepalive=(tk, tv)))
        op.finish()
        out = op.pull_device()
    cols = bench.device_cols(out, torch)
    g = cols[0][2]
    pids = torch.empty(g, dtype=torch.int32, device="cuda")
    ks, ka = native.ArrowSchema(), native.ArrowDeviceArray()
    bench._key_struct(native, ks, ka, cols[0][0], rank)
    native.check(native.lib.b200q_murmur3_partition(C.addressof(ks), C.addressof(ka), world, pids.data_ptr(), None))
    torch.cuda.synchronize()
    recv = exchange_columns([t for t, _, _ in cols], pids, world, dist)
    native.release_device_array(out)
    m = recv[0].numel()
    with native.NativeOp(plans["final_col"], conf, rank) as op:
        op.push_device(native.DeviceBatch([(t.data_ptr(), 0, m) for t in recv], m, rank, keepalive=recv))
        op.finish()
        res = op.pull_device()
        rc = bench.device_cols(res, torch)
        outk, outs, outc = (t.cpu().numpy().copy() for t, _, _ in rc)
        native.release_device_array(res)
    q.put((rank, outk, outs, outc, k, v))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_partial_exchange_final():
    from blaze_b200 import native
    if native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import cpu_ref
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    n = 400_000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    k = np.concatenate([r[4] for r in res]); v = np.concatenate([r[5] for r in res])
    ref = cpu_ref.hashagg_sum_count(k, v, nthreads=4)
    gk = np.concatenate([r[1] for r in res]); gs = np.concatenate([r[2] for r in res]); gc = np.concatenate([r[3] for r in res])
    assert len(np.unique(gk)) == len(gk), "every group must be owned by exactly one rank"
    og, orr = np.argsort(gk), np.argsort(ref["k"])
    assert np.array_equal(gk[og], ref["k"][orr]) and np.array_equal(gs[og], ref["sum"][orr]) and np.array_equal(gc[og], ref["count"][orr])
