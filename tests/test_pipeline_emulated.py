"""The whole pipeline on an EMULATED device, without a GPU: tools/emu/build_lib.sh builds the library (C ABI, plan decode,
stages, dispatch, every kernel) against a host stand-in for the CUDA runtime — one OS thread per CUDA thread — and
tools/emu/run_gpu_suite.py runs `-m gpu` parity tests on it.  Here: the committed golden fixtures (VM filter/project
with NULLs, Partial incl. the frozen Binary column, Final, the typed 2-key fused filter, f64/decimal128 aggregates).
The full GPU suite can be run the same way (`python tools/emu/run_gpu_suite.py tests -m gpu -k "not large_batch and not murmur3"`,
about an hour); it checks logic, never performance, and is no CPU path of the product."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu_dir(tmp_path_factory):
    """one build of the emulated library for the whole module (B200Q_EMU_REUSE)"""
    return str(tmp_path_factory.mktemp("emu"))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")
def test_golden_fixtures_on_the_emulated_device(emu_dir):
    env = dict(os.environ, B200Q_EMU_DIR=emu_dir, B200Q_EMU_REUSE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "run_gpu_suite.py"), os.path.join(ROOT, "tests", "test_golden_fixtures.py"),
                        "-m", "gpu", "-q", "-k", "not murmur3", "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=1800, cwd=ROOT)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")
def test_multi_rank_exchange_on_the_emulated_device(emu_dir):
    """N > 1 data path without GPUs: b200q_exchange_shuffle (murmur3 pids, counting partition, AllGather of the counts,
    AllToAllv) with 2 / 3 / 8 ranks run as threads over the NCCL stand-in of tools/emu (tests/test_gpu_exchange_threads.py)."""
    env = dict(os.environ, B200Q_EMU_DIR=emu_dir, B200Q_EMU_REUSE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "run_gpu_suite.py"), os.path.join(ROOT, "tests", "test_gpu_exchange_threads.py"),
                        "-m", "gpu", "-q", "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=1800, cwd=ROOT)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20)")
def test_shuffle_writer_on_the_emulated_device(emu_dir):
    """ShuffleWriterExec (pids + layout + tile-sort scatter/encode kernels, LZ4 framing, .data/.index) without a GPU: a subset of
    tests/test_gpu_shuffle_writer.py (the whole file passes the same way in ~4 min)."""
    env = dict(os.environ, B200Q_EMU_DIR=emu_dir, B200Q_EMU_REUSE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "run_gpu_suite.py"), os.path.join(ROOT, "tests", "test_gpu_shuffle_writer.py"),
                        "-m", "gpu", "-q", "-p", "no:cacheprovider", "-k", "chunks_can_stay or empty_input or outside_the_gpu_path or partial_aggregate"],
                       capture_output=True, text=True, env=env, timeout=1800, cwd=ROOT)
    assert r.returncode == 0 and "4 passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
