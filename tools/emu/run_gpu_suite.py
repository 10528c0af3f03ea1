"""Run (a subset of) the `-m gpu` parity tests against the EMULATED device: the whole library built by
tools/emu/build_lib.sh (host stand-in for the CUDA runtime, one OS thread per CUDA thread).  Checks the complete
pipeline logic — C ABI, plan decode, stages, kernel dispatch, every kernel — on a CPU-only box; it says nothing
about performance or the device memory model, and it is never used by the product (blaze_b200.native loads
blaze_b200/libblaze_b200.so; this runner swaps the handle explicitly, for the test process only).

usage: python tools/emu/run_gpu_suite.py [pytest args...]      e.g.  tests/test_gpu_agg.py -k "not large_batch" -x -q
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main(argv):
    out = os.environ.get("B200Q_EMU_DIR") or os.path.join(os.environ.get("TMPDIR", "/tmp"), "b200q_emu")
    lib = os.path.join(out, "libblaze_b200_emu.so")
    if not (os.environ.get("B200Q_EMU_REUSE") and os.path.exists(lib)):               # B200Q_EMU_REUSE=1: several runs share one build (tests/test_pipeline_emulated.py)
        lib = subprocess.run([os.path.join(HERE, "build_lib.sh"), out], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    from blaze_b200 import native
    native.LIB_PATH = lib
    native.lib = native._load()
    assert native.device_count() == 1
    import pytest
    return pytest.main(list(argv) or ["tests", "-m", "gpu", "-x", "-q"])


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
