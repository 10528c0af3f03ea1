"""CPU restatement of the reference's SortExec — TEST INFRASTRUCTURE ONLY.

datafusion-ext-plans/src/sort_exec.rs: rows are ordered by the memcmp order of their arrow-row encoded keys (:626-651):
per key column, in order, a NULL sorts before (nulls_first) or after every value, values ascend or descend, floats by IEEE
totalOrder (-NaN < -inf < ... < -0.0 < +0.0 < ... < +inf < +NaN); `fetch` keeps the first rows (:650, 946-960).  The
order among rows with equal keys is unspecified (unstable sort for short keys :637-643).
The row encoding itself is arrow-rs 55.2 (`arrow-row`, un-vendored): restated from its documented order.
Pinned by the reference's golden test_sort_i32 (sort_exec.rs:1447-1476) and by the property its fuzz test checks
(:1527-1607: same rows, ordered) — tests/test_sort_oracle.py.
"""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

from blaze_b200 import types as T
from oracle import blaze_oracle as O
from oracle.blaze_oracle import Batch, Col


def _order_word(dt, v):
    if dt.id == T.FLOAT64:
        b = struct.unpack("<Q", struct.pack("<d", float(v)))[0]
        return (~b & (2**64 - 1)) if b >> 63 else (b | 1 << 63)
    if dt.id == T.FLOAT32:
        b = struct.unpack("<I", struct.pack("<f", float(v)))[0]
        return (~b & (2**32 - 1)) if b >> 31 else (b | 1 << 31)
    return int(v)


def sort_key(cols: Sequence[Col], r: int, exprs: Sequence[Tuple[int, bool, bool]]):
    """a python tuple whose natural order is the row-encoding's memcmp order; exprs = [(column, descending, nulls_first)]"""
    k = []
    for ci, desc, nulls_first in exprs:
        c = cols[ci]
        if not c.valid[r]:
            k.append((0 if nulls_first else 2, 0))
        else:
            w = _order_word(c.dtype, c.values[r])
            k.append((1, -w if desc else w))
    return tuple(k)


def sort_exec(batches: Sequence[Batch], exprs: Sequence[Tuple[int, bool, bool]], fetch: Optional[int] = None) -> Optional[Batch]:
    if not batches:
        return None
    whole = O.concat_batches(batches[0].schema, list(batches))
    order = sorted(range(whole.num_rows), key=lambda r: sort_key(whole.cols, r, exprs))      # stable: ties keep arrival order
    if fetch is not None:
        order = order[:fetch]
    return whole.take(np.array(order, np.int64))
