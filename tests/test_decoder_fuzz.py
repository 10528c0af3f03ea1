"""Bounded sanitizer fuzz of the host-side decoders (protobuf plan reader, Arrow-IPC literal reader): hostile bytes at
the C-ABI boundary must end in a status code, never in a crash or undefined behaviour.  CPU only."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ with ASan/UBSan")
def test_plan_and_ipc_decoders_survive_mutated_inputs(tmp_path):
    env = dict(os.environ, TMPDIR=str(tmp_path))
    r = subprocess.run([os.path.join(ROOT, "tools", "fuzz", "run.sh"), "30000", "7"], capture_output=True, text=True, env=env, timeout=600)
    if r.returncode != 0 and ("cannot find -lasan" in r.stderr or "cannot find -lubsan" in r.stderr):
        pytest.skip("sanitizer runtimes are not installed")
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("ok=")]
    assert len(lines) == 2 and all(int(l.split()[0][3:]) > 0 for l in lines), r.stdout        # some mutants still decode: the harness really runs
