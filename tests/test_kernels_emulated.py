"""Logic tests of the specialised HashAgg update kernels WITHOUT a GPU: tools/emu compiles host-translated copies of
the kernel sources (inline PTX mapped onto host atomics, `__shared__` onto statics) and runs one OS thread per CUDA
thread with real warp/block rendezvous for shuffles, ballots and barriers.  Every kernel form (compacted-probe hash,
shared-memory dense, one-row-per-lane dense, gang dense, experimental hot-key cache) x {1, 2 keys} x {lean, typed +
NULLs} x {no filter, fused filter} is compared group by group with a plain host loop — once with the kernel chosen by the test and once through the
library's own dispatcher (`launch_agg_fast_update`); the key-range and skew-probe launchers are checked too.  It checks the kernels'
LOGIC (lane exchange, layouts, insert protocol, fall-back of out-of-range keys) — not timing, not the memory model
of the device; parity on hardware is the job of the `-m gpu` tests."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")
def test_hashagg_kernel_forms_under_emulation(tmp_path):
    env = dict(os.environ, TMPDIR=str(tmp_path))
    # the CPU suite runs the typed + NULLs half of the configurations (every kernel form, filter shape and key count is in it);
    # `tools/emu/run.sh` without an argument runs all of them (~3 min)
    r = subprocess.run([os.path.join(ROOT, "tools", "emu", "run.sh"), "typed"], capture_output=True, text=True, env=env, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0, tail + r.stderr[-2000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.endswith("0 failed") and int(last.split()[0]) >= 100, last
