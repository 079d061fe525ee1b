"""Host-only C++ under AddressSanitizer + UBSan: tests/c_abi/sanitized_host.cpp drives the planner (build_plan with
the default options: edge attachment + augmentation, colouring + recolouring, rim merge, hierarchical leftovers) and
the ingest code on a few meshes.  compute-sanitizer is closed on the GPU pool; this covers the code that runs inside
sb_create / sb_plan / sb_tetmesh_* on the host."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "softbodyunity_b200", "csrc")


def _build(tmp_path_factory, kind):
    out = tmp_path_factory.mktemp("san") / ("sanitized_host_" + kind.split(",")[0])
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=" + kind] + (["-fno-sanitize-recover=undefined"] if "undefined" in kind else []) + [

           "-fno-omit-frame-pointer", "-ffp-contract=off", "-march=x86-64-v3", "-pthread", "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "sanitized_host.cpp"), os.path.join(CSRC, "plan.cpp"), os.path.join(CSRC, "ingest.cpp"),
           "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("no sanitizer runtime on this machine")
    assert r.returncode == 0, r.stderr[-2000:]
    return str(out)


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    return _build(tmp_path_factory, "address,undefined")


@pytest.fixture(scope="module")
def exe_tsan(tmp_path_factory):
    return _build(tmp_path_factory, "thread")


@pytest.mark.parametrize("args", ["20 20 20 0", "30 24 18 300", "14 12 11 256", "8 8 8 64", "33 9 40 0"])
def test_planner_and_ingest_are_clean_under_asan_ubsan(exe, args, tmp_path):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe] + args.split() + [str(tmp_path)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr, (r.stdout + r.stderr)[-3000:]
    assert "ingest V" in r.stdout and "load 0 0" in r.stdout


def test_planner_threads_are_race_free_under_tsan(exe_tsan, tmp_path):
    # the planner builds the tiles of a pass on several threads (4 here)
    r = subprocess.run([exe_tsan, "30", "24", "18", "300", str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr, (r.stdout + r.stderr)[-3000:]
