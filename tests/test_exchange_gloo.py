"""world_size-2/4/8 gloo tests of the partial->final exchange plumbing (CPU): partition ids follow the
reference's murmur3(seed 42) pmod rule (oracle), every group ends up on exactly one owner rank, and
merging what each rank received reproduces the single-process aggregate."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from exchange_reference import exchange_columns
    from blaze_b200 import types as T
    from oracle import blaze_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = 5000
    k = rng.integers(0, 300, n, dtype=np.int64); v = rng.integers(-1000, 1000, n, dtype=np.int64)
    # this rank's partial aggregate (Partial mode): one state row per local group
    keys, inv = np.unique(k, return_inverse=True)
    sums = np.zeros(len(keys), np.int64); np.add.at(sums, inv, v)
    cnts = np.bincount(inv, minlength=len(keys)).astype(np.int64)
    hashes = O.create_murmur3_hashes([O.Col(T.int64, keys, np.ones(len(keys), bool))], len(keys), 42)
    pids = torch.from_numpy(O.partition_ids(hashes, world).astype(np.int64))
    rk, rs, rc = exchange_columns([torch.from_numpy(keys), torch.from_numpy(sums), torch.from_numpy(cnts)], pids, world, dist)
    # Final mode on the owner: merge the received partial states
    fk, finv = np.unique(rk.numpy(), return_inverse=True)
    fs = np.zeros(len(fk), np.int64); np.add.at(fs, finv, rs.numpy())
    fc = np.zeros(len(fk), np.int64); np.add.at(fc, finv, rc.numpy())
    owner_ok = bool(np.all(O.partition_ids(O.create_murmur3_hashes([O.Col(T.int64, fk, np.ones(len(fk), bool))], len(fk), 42), world) == rank))
    q.put((rank, fk.tolist(), fs.tolist(), fc.tolist(), k.tolist(), v.tolist(), owner_ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 8 + world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_k = np.concatenate([np.array(r[4]) for r in res]); all_v = np.concatenate([np.array(r[5]) for r in res])
    exp = {}
    for kk, vv in zip(all_k, all_v):
        s, c = exp.get(kk, (0, 0)); exp[kk] = (s + vv, c + 1)
    got = {}
    for _, fk, fs, fc, _, _, owner_ok in res:
        assert owner_ok
        for kk, ss, cc in zip(fk, fs, fc):
            assert kk not in got, "a group must have exactly one owner"
            got[kk] = (ss, cc)
    assert got == exp
