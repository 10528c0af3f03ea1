"""N-GPU path (when the box has >= 2 GPUs): Partial per GPU -> b200q_exchange_shuffle (murmur3(42) pmod N ownership,
in-library NCCL AllToAllv of the columnar partial states) -> Final per GPU; the union of the ranks' results must equal
the single-process aggregate computed by the C restatement of the reference algorithm, every group on exactly one rank
and on the rank the reference's partitioning rule (shuffle/mod.rs:163-188) names."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from blaze_b200 import native
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(native.exchange_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ex = native.Exchange(uid.cpu().numpy().tobytes(), rank, world, rank)
    rng = np.random.default_rng(7 + rank)
    k = rng.integers(0, 50_000, n, dtype=np.int64); v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
    tk, tv = torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda()
    plans = bench.build_plans("M1")
    conf = native.default_conf(partial_state_columnar=1)
    with native.NativeOp(plans["partial_col"], conf, rank) as op:
        op.push_device(native.DeviceBatch([(tk.data_ptr(), 0, n), (tv.data_ptr(), 0, n)], n, rank, keepalive=(tk, tv)))
        op.finish()
        out = op.pull_device()
        schema = op.output_schema()
    recv = ex.shuffle(schema, out, 1)
    assert ex.kernel_launches() > 0
    with native.NativeOp(plans["final_col"], conf, rank) as op:
        op.push_device_array(recv)
        op.finish()
        res = op.pull_device()
        outk, outs, outc = (t.cpu().numpy().copy() for t in bench.device_cols(res, torch))
        native.release_device_array(res)
    q.put((rank, outk, outs, outc, k, v))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_partial_exchange_final_across_gpus(world):
    from blaze_b200 import native, types as T
    if native.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from oracle import blaze_oracle as O, cpu_ref
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    n = 400_000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
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
    for rank, outk, _, _, _, _ in res:
        pid = O.partition_ids(O.create_murmur3_hashes([O.Col(T.int64, outk, np.ones(len(outk), bool))], len(outk), 42), world)
        assert np.all(pid == rank), "group ownership must follow pmod(murmur3(key, 42), world)"
