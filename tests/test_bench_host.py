"""bench.py on the host: the byte model of SURVEY.md 8(d), the reference arm's JSON line (it runs on the CPU by
construction) and that both arms describe a run by the same `config` object.  The native arm needs a GPU; bench.py
refuses to run without one."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_per_vertex_substep_at_the_headline_mesh():
    # 64 + 10 * (12 E/V + 20 T/V + 32) + 64 = 2127.5 B at 100^3
    assert bench.bytes_per_substep(1_000_000, 5_910_300, 4_851_495, 10) == pytest.approx(2127.535, abs=1e-3)
    assert bench.bytes_per_substep(8_000_000, 47_640_600, 39_402_995, 10) == pytest.approx(2147.7, abs=0.1)


def _run(*argv, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, **(env or {})))
    return r


def test_reference_arm_prints_the_contract_line_and_the_native_arms_config():
    r = _run("--impl", "reference", "--workload", "block", "--n", "14", "--steps", "2", "--warmup", "3", "--tile-cap", "256")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "vertex-substeps/sec" and line["unit"] == "vertex-substeps/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 3 and line["n_gpus"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] == (os.cpu_count() or 1) and "sample" in cb
    assert line["value"] == pytest.approx(14 ** 3 * 2 / (line["ms_per_step"] * 2e-3), rel=1e-6)
    # the same object the native arm prints for these arguments (describe_config over the same plan)
    from softbodyunity_b200 import SoftBody
    args = bench.parse_args(["--workload", "block", "--n", "14", "--tile-cap", "256"])
    pos, tets, tris, name = bench.workload(args)
    plan = SoftBody(pos, tets, tris, host_only=True, substeps=args.substeps, iterations=args.iterations, flags=bench.solver_flags(args),
                    **bench.plan_options(args))
    assert line["config"] == json.loads(json.dumps(bench.describe_config(args, plan.info(), name, len(pos), 1)))
    assert line["config"]["workload"].startswith("block14^3 (2744 verts)")


def test_other_ranks_of_the_reference_arm_exit_without_work():
    r = _run("--impl", "reference", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--workload", "block", "--n", "8", "--steps", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)


def test_the_plan_options_of_the_distributed_workload_do_not_depend_on_the_rank_count():
    a = bench.parse_args(["--workload", "dist"])
    assert bench.plan_options(a)["dist_ranks"] == 8 and bench.plan_options(a)["block_threads"] == 160
    a = bench.parse_args(["--workload", "dist", "--slabs", "--block-threads", "128"])
    assert bench.plan_options(a)["dist_ranks"] == 0 and bench.plan_options(a)["block_threads"] == 128


def test_a_traffic_capture_of_another_plan_is_refused():
    info = {"n_tile_passes": 2, "rounds_in_pass": [3, 4], "tiles_in_pass": [5, 6]}
    t, why = bench.measured_traffic(info)
    assert t is None and "another plan" in why


def test_the_oracle_checksum_fixture_is_found_by_the_key_bench_uses(tmp_path, monkeypatch):
    # tests/golden/make_dist_checksum.py writes {key: {"after_frames": {frames: checksum}}}; bench.py looks the GPUs' run up
    # by the same key (mesh, stepping, plan) and reports `matches_cpu_oracle`
    out = tmp_path / "tests" / "golden"
    out.mkdir(parents=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_dist_checksum.py"), "--n", "20", "--frames", "1", "2",
                        "--out", str(out / "dist_checksum.json")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    doc = json.load(open(out / "dist_checksum.json"))
    (key, entry), = doc.items()
    assert key.startswith("dist n=20 S=10 I=10 | tiles [") and set(entry["after_frames"]) == {"1", "2"}
    assert entry["after_frames"]["1"] != entry["after_frames"]["2"]
    from softbodyunity_b200 import SoftBody
    args = bench.parse_args(["--workload", "dist", "--n", "20"])
    pos, tets, tris, name = bench.workload(args)
    plan = SoftBody(pos, tets, tris, host_only=True, substeps=10, iterations=10, flags=0, **bench.plan_options(args))
    cfg = bench.describe_config(args, plan.info(), name, len(pos), 2)   # any rank count: the key does not depend on it
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.oracle_checksum(cfg, args, 2) == entry["after_frames"]["2"]
    assert bench.oracle_checksum(cfg, args, 3) is None
    # the committed fixture, if present, is well-formed
    monkeypatch.undo()
    path = os.path.join(ROOT, "tests", "golden", "dist_checksum.json")
    if os.path.exists(path):
        for k, e in json.load(open(path)).items():
            assert " | tiles [" in k and all(len(v.split("-")) == 2 for v in e["after_frames"].values())


def test_reference_arm_on_rank_0_of_two_prints_the_distributed_workloads_config():
    # N > 1: the default workload is the mesh that shards; rank 0 alone runs the CPU arm and names the run as the native arm does
    r = _run("--impl", "reference", "--gpus", "2", "--size", "20", "--steps", "1", "--warmup", "3",
             env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    from softbodyunity_b200 import SoftBody
    args = bench.parse_args(["--workload", "dist", "--size", "20"])
    pos, tets, tris, name = bench.workload(args, 0, 2)
    assert "over 2 rank(s)" in name and line["config"]["workload"] == name
    plan = SoftBody(pos, tets, tris, host_only=True, substeps=10, iterations=10, flags=0, **bench.plan_options(args))
    assert line["config"] == json.loads(json.dumps(bench.describe_config(args, plan.info(), name, len(pos), 2)))
    assert "comm" in line["config"] and "partition" in line["config"]


CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline")


def test_native_arm_plumbing_on_one_rank_with_a_stand_in_solver():
    # tests/bench_mock.py: no GPU, made-up timings -- only the shape of the line and the path through the script are checked
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_mock.py"), "--workload", "block", "--size", "20", "--steps", "3",
                        "--tile-cap", "256", "--bodies-n", "4", "--kernel-breakdown"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert all(k in d for k in CONTRACT) and d["n_gpus"] == 1 and d["scaling"] == "strong"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "protocol"}
    assert d["e2e"]["h2d_bytes_per_step"] == 32 * 8000 and d["e2e_component"]["h2d_bytes_per_step"] == 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "launch_ms", "step_frac", "step_frac_of_nominal_8000"}
    assert d["roofline"]["traffic"] is None and "another plan" in d["roofline"]["traffic_source"]   # the capture is of the 1 M plan
    assert d["cpu_baseline"]["kind"] == "port" and d["bodies"]["n_verts"] == 4 * 2028
    assert d["state_checksum"]["after_frames"] == 6 and d["state_checksum"]["matches_cpu_oracle"] is None
    assert "NOT A BENCH CONFIGURATION" in d["config"]["l2"] and "kernel_breakdown_ms" in d


def test_native_arm_plumbing_on_two_gloo_ranks_with_a_stand_in_solver():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29788", os.path.join(ROOT, "tests", "bench_mock.py"), "--gpus", "2", "--size", "24", "--steps", "3",
                        "--bodies-n", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                   # rank 0 alone prints
    d = json.loads(lines[0])
    assert all(k in d for k in CONTRACT) and d["n_gpus"] == 2 and d["scaling"] == "strong"
    assert "single block24^3 (13824 verts) over 2 rank(s)" in d["config"]["workload"] and d["config"]["n_verts"] == 13824
    assert "comm" in d["config"] and d["config"]["partition"].startswith("compact blocks")
    assert d["value"] == pytest.approx(13824 * 10 / (d["ms_per_step"] * 1e-3))        # the WHOLE mesh over the max-over-ranks time
    assert d["e2e"]["h2d_bytes_per_step"] == 32 * 13824                              # summed over the ranks: every vertex once
    assert d["distribution"]["own_verts_rank0"] == 13824 // 2 and len(d["distribution"]["tiles_rank0"]) == d["config"]["tile_passes"]
    assert d["roofline"]["peak"] == pytest.approx(2 * bench.peaks()[0]) and d["cpu_baseline"] is None
    assert d["state_checksum"]["x4_words_hi_lo"] == "%x-%x" % (13824 * 4 * 0x3f80, 0)  # all-ones state of the stand-in, all ranks
    assert d["bodies"]["n_verts"] == 4 * 2028 and d["bodies"]["bodies_rank0"] == 2
