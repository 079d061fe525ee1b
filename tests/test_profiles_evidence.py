"""The committed measurement evidence is self-consistent: the bench line of the round's final state carries every key of
the contract, its roofline arithmetic adds up, the ncu traffic capture is of the plan that was benched, and the state
checksum the 8-GPU run of the 8 M-vertex mesh printed is the CPU oracle's (tests/golden/dist_checksum.json)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_final_bench_line_has_the_contract_keys_and_consistent_arithmetic():
    d = _line(os.path.join(P, "bench_r02z.json"))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "vertex-substeps/sec" and d["dtype"] == "f32" and d["n_gpus"] == 1 and d["warmup"] >= 3
    c = d["config"]
    assert c["n_verts"] == 1_000_000 and "block100^3" in c["workload"]
    # value = V * substeps / time per step
    assert d["value"] == pytest.approx(c["n_verts"] * 10 / (d["ms_per_step"] * 1e-3), rel=1e-9)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    # the tile launches carry the whole algorithmic traffic of the step (no separate per-vertex kernels)
    assert r["separate_vertex_kernels_per_step"] == 0
    assert r["bytes_per_launch"] * r["launches_per_step"] == pytest.approx(r["bytes_per_vertex_substep"] * c["n_verts"] * 10, rel=1e-9)
    assert r["achieved"] == pytest.approx(r["bytes_per_launch"] / (r["launch_ms"] * 1e-3) / 1e9, rel=1e-9)
    assert r["step_achieved"] == pytest.approx(d["value"] * r["bytes_per_vertex_substep"] / 1e9, rel=1e-9)
    assert 0.5 < r["frac"] < 0.6 and r["step_frac"] <= r["frac"] + 1e-3
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 32 * c["n_verts"] and e["d2h_bytes_per_step"] > e["h2d_bytes_per_step"]
    assert e["value"] < d["value"]                     # copies inside the timed region
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] == d["clocks"]["sm_max_mhz"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == (r["launches_per_step"] + 1) * d["steps"]   # + the normals launch


def test_the_traffic_capture_is_of_the_benched_plan():
    d = _line(os.path.join(P, "bench_r02z.json"))
    t = json.load(open(os.path.join(P, "r02z_traffic.json")))
    assert t["rounds_per_sweep"] == d["config"]["rounds_per_sweep"] and t["tiles_in_pass"] == d["config"]["tiles_in_pass"]
    per_launch = [x["dram_read_bytes"] + x["dram_write_bytes"] for x in t["launches"]]
    # real DRAM traffic per launch stays below the algorithmic bytes per launch: no wasted re-reads
    assert max(per_launch) < d["roofline"]["bytes_per_launch"]


def test_the_eight_gpu_run_printed_the_oracles_checksum():
    d = _line(os.path.join(P, "r02", "bench_r02p_8M_n8.json"))
    assert d["n_gpus"] == 8 and d["config"]["n_verts"] == 8_000_000 and d["scaling"] == "strong"
    chk = d["state_checksum"]
    fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "dist_checksum.json")))
    (entry,) = [v for k, v in fixture.items() if k.startswith("dist n=200 ") and str(d["config"]["tiles_in_pass"]) in k
                and "rounds %d " % d["config"]["rounds_per_sweep"] in k and "bt %d " % d["config"]["block_threads"] in k]
    assert entry["after_frames"][str(chk["after_frames"])] == chk["x4_words_hi_lo"]
    # 8 GPUs on the 8 M mesh against the single-GPU headline figure: the >= 6x of BASELINE.json
    one = _line(os.path.join(P, "bench_r02z.json"))
    assert d["value"] / one["value"] > 6.0
