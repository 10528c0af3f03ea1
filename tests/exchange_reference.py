"""TEST-ONLY reference of the multi-GPU partial->final exchange contract, on torch.distributed (gloo on
CPU): rows travel to rank pmod(murmur3(keys, 42), world) (shuffle/mod.rs:163-188) and arrive grouped by source rank.
The product path is b200q_exchange_shuffle in blaze_b200/csrc/exchange.cu (device-side partition + NCCL AllToAllv);
this file only lets tests/test_exchange_gloo.py exercise world_size 2/4/8 ownership + merge logic without GPUs."""
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
