#!/usr/bin/env python
"""bench.py — rows/sec of the hot path on synthetic TPC-DS-shaped batches (BASELINE.json).

Workload (config.workload = "M1", BASELINE.json configs[1]): HashAggregateExec SUM(v), COUNT(v)
GROUP BY k over `rows` int64/int64 rows per GPU, k ~ U[0, 2^20) (1M groups), v ~ U[-1e6, 1e6)
(models UnscaledValue(decimal(7,2))).  A step = one complete aggregation of the batch:
Partial -> (exchange when N > 1) -> Final, results pulled.

  value      whole-job rows/s with the input already resident in HBM (push_device)
  e2e        the same through the host-buffer C ABI (b200q_op_push of pinned host Arrow batches, result
             pulled back to the host): H2D/D2H inside the timed region
  roofline   HBM: algorithmic 16 B/row (+24 B/group out) / CUDA-event time of the update kernel
  cpu_baseline  oracle/cpu_ref.c (restatement of the reference CPU algorithm) on this box's host cores

`--impl reference` times that CPU restatement alone (the reference binary cannot be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CARD = 1 << 20
ALG_BYTES_PER_ROW = 16.0          # read k,v once
ALG_BYTES_PER_GROUP = 24.0        # key + sum + count written once


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=10)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def m1_plans():
    from blaze_b200 import exprs as E, plans as PL, types as T
    ins = T.Schema([T.Field("k", T.int64, False), T.Field("v", T.int64, False)])
    leaf = PL.MemoryExec(ins)
    g = [E.GroupingExpr("k", E.Column("k"))]
    mk = lambda mode, ch: [E.AggExpr("sum_v", mode, PL.create_agg(E.AGG_SUM, ch, ins, T.int64)),
                           E.AggExpr("count_v", mode, PL.create_agg(E.AGG_COUNT, ch, ins, T.int64))]
    partial = PL.AggExec(PL.HashAgg, g, mk(E.PARTIAL, [E.Column("v")]), True, leaf)
    final = PL.AggExec(PL.HashAgg, g, mk(E.FINAL, [E.placeholder(T.int64)]), False, partial)
    partial_col = PL.AggExec(PL.HashAgg, g, mk(E.PARTIAL, [E.Column("v")]), True, leaf, columnar_state=True)
    final_leaf = PL.MemoryExec(partial_col.schema())
    final_col = PL.AggExec(PL.HashAgg, g, mk(E.FINAL, [E.placeholder(T.int64)]), False, final_leaf)
    return dict(single=final.plan_bytes(), partial_col=partial_col.plan_bytes(), final_col=final_col.plan_bytes())


class CudaView:
    """zero-copy torch view of a device buffer returned by pull_device"""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
        self.owner = owner


def device_cols(dev_array, torch):
    """ArrowDeviceArray (struct of fixed-width int64 columns) -> [(values int64 tensor, validity ptr)]"""
    a = dev_array.array
    out = []
    for i in range(a.n_children):
        c = a.children[i].contents
        vptr = c.buffers[1]
        t = torch.as_tensor(CudaView(vptr, c.length * 8, dev_array), device="cuda").view(torch.int64)
        out.append((t, c.buffers[0], c.length))
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from blaze_b200 import native
    from blaze_b200.exchange import exchange_columns
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rows = env_int("B200Q_BENCH_ROWS", 1_000_000_000)
    e2e_batch = env_int("B200Q_BENCH_E2E_BATCH_ROWS", 1 << 24)
    # the e2e leg keeps its input in PINNED host memory (16 B/row): all ranks of the node together stay below
    # min(64 GB, 35 % of MemAvailable) so that an 8-rank run cannot drive the box out of memory
    e2e_rows = env_int("B200Q_BENCH_E2E_ROWS", 0)
    if e2e_rows <= 0:
        budget = min(64e9, 0.35 * mem_available_bytes())
        e2e_rows = int(min(rows, max(e2e_batch, budget / (16 * world) // e2e_batch * e2e_batch)))
    plans = m1_plans()
    gen = torch.Generator(device=dev); gen.manual_seed(44 + rank)
    k = torch.randint(0, CARD, (rows,), dtype=torch.int64, device=dev, generator=gen)
    v = torch.randint(-10**6, 10**6, (rows,), dtype=torch.int64, device=dev, generator=gen)
    torch.cuda.synchronize()
    conf = native.default_conf(agg_initial_groups=CARD)
    conf_col = native.default_conf(agg_initial_groups=CARD, partial_state_columnar=1)
    stats = {"launches": 0, "hot_ns": 0, "hot_rows": 0, "hot_launches": 0, "groups": 0}

    def acc_metrics(m):
        stats["launches"] += m["gpu_kernel_launches"]; stats["hot_ns"] += m["hot_kernel_ns"]
        stats["hot_rows"] += m["hot_kernel_rows"]; stats["hot_launches"] += m["hot_kernel_launches"]

    def step_device():
        if world == 1:
            with native.NativeOp(plans["single"], conf, local) as op:
                op.push_device(native.DeviceBatch([(k.data_ptr(), 0, rows), (v.data_ptr(), 0, rows)], rows, local, keepalive=(k, v)))
                op.finish()
                out = op.pull_device()
                stats["groups"] = out.array.length
                native.release_device_array(out)
                acc_metrics(op.metrics())
            return
        # N > 1: Partial per GPU -> murmur3(seed 42) pmod N ownership -> all_to_all of partial states -> Final per GPU
        with native.NativeOp(plans["partial_col"], conf_col, local) as op:
            op.push_device(native.DeviceBatch([(k.data_ptr(), 0, rows), (v.data_ptr(), 0, rows)], rows, local, keepalive=(k, v)))
            op.finish()
            out = op.pull_device()
            acc_metrics(op.metrics())
        cols = device_cols(out, torch)
        g = cols[0][2]
        pids = torch.empty(g, dtype=torch.int32, device=dev)
        ks = native.ArrowSchema(); ka = native.ArrowDeviceArray()
        _key_struct(native, ks, ka, cols[0][0], local)
        native.check(native.lib.b200q_murmur3_partition(native.C.addressof(ks), native.C.addressof(ka), world, pids.data_ptr(), None))
        torch.cuda.synchronize()
        recv = exchange_columns([t for t, _, _ in cols], pids, world, dist)
        native.release_device_array(out)
        n_in = recv[0].numel()
        with native.NativeOp(plans["final_col"], conf_col, local) as op:
            op.push_device(native.DeviceBatch([(t.data_ptr(), 0, n_in) for t in recv], n_in, local, keepalive=recv))
            op.finish()
            res = op.pull_device()
            stats["groups"] = res.array.length if res is not None else 0
            if res is not None:
                native.release_device_array(res)
            acc_metrics(op.metrics())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    for key in stats:
        stats[key] = 0 if key != "groups" else stats[key]
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = tmax.item() / args.steps
    value = rows * world / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------------
    import pyarrow as pa
    numa = bind_to_gpu_numa(torch, local)        # pinned buffers on the GPU's NUMA node (first touch)
    hk = torch.empty(e2e_rows, dtype=torch.int64, pin_memory=True); hv = torch.empty(e2e_rows, dtype=torch.int64, pin_memory=True)
    hk.copy_(k[:e2e_rows]); hv.copy_(v[:e2e_rows]); torch.cuda.synchronize()

    def host_batches():
        out = []
        schema = pa.schema([pa.field("k", pa.int64(), False), pa.field("v", pa.int64(), False)])
        for b in range(0, e2e_rows, e2e_batch):
            m = min(e2e_batch, e2e_rows - b)
            arrs = [pa.Array.from_buffers(pa.int64(), m, [None, pa.foreign_buffer(t.data_ptr() + 8 * b, 8 * m, base=t)]) for t in (hk, hv)]
            out.append(pa.RecordBatch.from_arrays(arrs, schema=schema))
        return out
    hb = host_batches()
    e2e_stats = {"h2d": 0, "d2h": 0}

    def step_e2e():
        # single-GPU form of the public API path (for N > 1 every rank runs it on its own shard)
        with native.NativeOp(plans["single"], conf, local) as op:
            for b in hb:
                op.push(b)
            op.finish()
            n_out = 0
            while True:
                o = op.pull()
                if o is None:
                    break
                n_out += o.num_rows
            m = op.metrics()
            e2e_stats["h2d"], e2e_stats["d2h"] = m["h2d_bytes"], m["d2h_bytes"]
        return n_out

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        step_e2e()
    e1.record()
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = e2e_rows * world / (t_e2e.item() / e2e_steps * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))); peak_src = "measured"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    launches = max(1, stats["hot_launches"])
    alg_bytes_per_launch = (ALG_BYTES_PER_ROW * stats["hot_rows"] + ALG_BYTES_PER_GROUP * stats["groups"] * args.steps) / launches
    achieved = (ALG_BYTES_PER_ROW * stats["hot_rows"]) / max(1, stats["hot_ns"])        # bytes/ns == GB/s
    traffic = None
    try:   # DRAM bytes of one launch of the same kernel/size from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        rows_per_launch = stats["hot_rows"] / launches
        if abs(rows_per_launch - tj["rows_per_launch"]) / tj["rows_per_launch"] < 0.1:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "agg_update", "launches": stats["hot_launches"],
                "avg_launch_ms": stats["hot_ns"] / launches / 1e6, "alg_bytes_per_launch": alg_bytes_per_launch}
    # ---- CPU baseline (restatement of the reference algorithm) on this box's host cores --------------------
    cpu = None
    os.sched_setaffinity(0, range(os.cpu_count() or 1))
    if world == 1:
        cpu = cpu_baseline(hk.numpy(), hv.numpy(), min(e2e_rows, env_int("B200Q_BENCH_CPU_ROWS", 1 << 28)))
    line = {
        "metric": "rows/sec on TPC-DS q1 hash-agg+filter at 1/2/4/8 B200; HBM GB/s vs 8 TB/s", "value": value, "unit": "rows/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "M1: HashAggregateExec SUM(v),COUNT(v) GROUP BY k; k~U[0,2^20) int64, v~U[-1e6,1e6) int64 (BASELINE.json configs[1])",
                   "rows_per_gpu": rows, "groups": CARD, "parallelism": f"dp{world}" + ("" if world == 1 else " + murmur3 pmod all_to_all of partial states"),
                   "l2_policy": "input (%.1f GB/GPU) is far larger than the 126 MB L2; no flush needed" % (rows * 16 / 1e9), "plan": "AggExec(Partial) -> AggExec(Final), reference protobuf + C ABI"},
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": e2e_stats["h2d"], "d2h_bytes_per_step": e2e_stats["d2h"],
                "rows_per_gpu": e2e_rows, "host_batch_rows": e2e_batch, "steps": e2e_steps, "host_numa_node": numa},
        "gpu_launches": stats["launches"], "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 64e9


def bind_to_gpu_numa(torch, local):
    """run this process (and first-touch its pinned buffers) on the CPU socket the GPU hangs off"""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def _key_struct(native, ks, ka, key_tensor, device):
    """struct<k:int64> ArrowSchema + ArrowDeviceArray over a device tensor (for b200q_murmur3_partition)"""
    C = native.C
    child_s = native.ArrowSchema(); child_s.format = b"l"; child_s.name = b"k"; child_s.flags = 0
    ks.format = b"+s"; ks.name = b""; ks.n_children = 1
    arr = (C.POINTER(native.ArrowSchema) * 1)(C.pointer(child_s)); ks.children = C.cast(arr, C.POINTER(C.POINTER(native.ArrowSchema)))
    ks._keep = (child_s, arr)
    db = native.DeviceBatch([(key_tensor.data_ptr(), 0, key_tensor.numel())], key_tensor.numel(), device, keepalive=(key_tensor,))
    C.memmove(C.addressof(ka), C.addressof(db.dev), C.sizeof(native.ArrowDeviceArray))
    ka._keep = db
    native.DeviceBatch._live.pop(db._id, None)


def usable_cores():
    """host threads this container may actually run: min(visible CPUs, cgroup cpu.max quota).
    (oracle/cpu_ref.c peaks there: profiles/r01_cpu_ref_thread_scaling.txt)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def cpu_baseline(k_np, v_np, sample_rows, threads=None):
    from oracle import cpu_ref
    threads = threads or usable_cores()
    k, v = k_np[:sample_rows], v_np[:sample_rows]
    cpu_ref.hashagg_time_only(k[: min(sample_rows, 1 << 22)], v[: min(sample_rows, 1 << 22)], threads)    # warm-up
    t0 = time.perf_counter()
    cpu_ref.hashagg_time_only(k, v, threads)
    dt = time.perf_counter() - t0
    return {"value": sample_rows / dt, "unit": "rows/s", "cores": threads, "kind": "port",
            "sample": f"{sample_rows} rows of the same M1 batch, {threads} reference-style tasks + final merge (oracle/cpu_ref.c)", "seconds": dt}


def run_reference(args):
    """The reference arm: the restatement of the reference's own CPU algorithm on all host threads.
    (The Rust reference cannot be built or installed in this image: no cargo/rustc, no network.)"""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu_ref
    threads = usable_cores()
    rows = env_int("B200Q_BENCH_REF_ROWS", 1 << 28)
    rng = np.random.default_rng(44)
    k = rng.integers(0, CARD, rows, dtype=np.int64); v = rng.integers(-10**6, 10**6, rows, dtype=np.int64)
    for _ in range(args.warmup):
        cpu_ref.hashagg_time_only(k, v, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_ref.hashagg_time_only(k, v, threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = rows / dt
    print(json.dumps({
        "impl": "reference", "metric": "rows/sec on TPC-DS q1 hash-agg+filter at 1/2/4/8 B200; HBM GB/s vs 8 TB/s", "value": val, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "M1: HashAggregateExec SUM(v),COUNT(v) GROUP BY k; k~U[0,2^20) int64, v~U[-1e6,1e6) int64 (BASELINE.json configs[1])",
                   "rows_per_step": rows, "groups": CARD},
        "cpu_baseline": {"value": val, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": f"{rows} rows per step, {threads} reference-style tasks + final merge (oracle/cpu_ref.c)"},
        "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
