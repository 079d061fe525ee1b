"""Real GPUs (skipped on a box with fewer than two): the peer-memory distribution under torchrun must reproduce the CPU
oracle bit for bit -- state and surface normals -- on 2 and 4 GPUs.  Virtual ranks on one stream (tests/test_dist.py)
cannot show a cross-GPU memory-ordering bug; this can.  The logs of the round's runs are kept under profiles/."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(n, script, *args, port=29541):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.mark.gpu
@pytest.mark.parametrize("n,extra", [(2, []), (2, ["--iterations", "5", "--flags", "16"]), (2, ["--slabs"]), (2, ["--flags", "8"]),
                                     (4, []), (8, [])])  # flags 16: no PDL; 8: no normals (the frame ends with k_dist_sync)
def test_peer_memory_distribution_on_real_gpus_is_bit_identical(n, extra):
    if _gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    r = _torchrun(n, "mp_run_dist.py", "--dims", "40", "40", "80", "--frames", "4", "--substeps", "5", *extra, port=29541 + n)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-identical to the CPU oracle (state and surface normals): True" in r.stdout, r.stdout[-2000:]
    assert "peer wait timed out: False" in r.stdout


@pytest.mark.gpu
def test_partitioned_ghost_scheme_on_real_gpus_is_bit_identical():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "mp_run_partitioned.py", "--frames", "4", port=29551)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-identical to the CPU oracle (combined order): True" in r.stdout, r.stdout[-2000:]
