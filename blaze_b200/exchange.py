"""Multi-GPU partial->final exchange of aggregate states.

The reference repartitions partial-aggregate rows with Spark's shuffle: partition id =
pmod(murmur3_x86_32(key columns, seed 42), n) (datafusion-ext-plans/src/shuffle/mod.rs:163-188) and the
Final AggExec of each reduce partition merges what it receives.  Here every rank owns one partition and
the rows travel through one NCCL AllToAllv per state column over NVLink (torch.distributed plumbing);
the partition ids come from the library's murmur3 kernel, so GPU partitions equal Spark reduce partitions.

`exchange_columns` is backend-agnostic (gloo on CPU in tests/test_exchange_gloo.py, nccl on GPUs).
"""
from __future__ import annotations

from typing import List, Sequence


def exchange_columns(cols: Sequence, pids, world: int, dist) -> List:
    """cols: equally long 1-D torch tensors (the key + state columns of this rank's partial result);
    pids: int tensor with the owner rank of every row.  Returns the columns of the rows this rank owns
    (concatenated in source-rank order)."""
    import torch
    order = torch.argsort(pids, stable=True)
    send_counts = torch.bincount(pids.to(torch.int64), minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    out = []
    for t in cols:
        src = t[order].contiguous()
        dst = torch.empty(sum(rc), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(dst, src, output_split_sizes=rc, input_split_sizes=sc)
        out.append(dst)
    return out
