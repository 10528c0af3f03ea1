"""thread scaling of the CPU restatement on this box (decides cpu_baseline.cores)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cpu_ref
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
rng = np.random.default_rng(44); n = 1 << 27
k = rng.integers(0, 1 << 20, n, dtype=np.int64); v = rng.integers(-10**6, 10**6, n, dtype=np.int64)
for t in (1, 4, 8, 16, 32, 64, 128):
    t0 = time.perf_counter(); g = cpu_ref.hashagg_time_only(k, v, t); dt = time.perf_counter() - t0
    print(f"threads={t:4d} groups={g} {n/dt/1e6:9.2f} Mrows/s  {dt:.2f}s", flush=True)
