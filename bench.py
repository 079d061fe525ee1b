#!/usr/bin/env python
"""bench.py -- vertex-substeps/sec of the soft-body substep path (BASELINE.json metric).

A "step" is one frame of the workload: `substeps` substeps, each with `iterations`
projection sweeps, over the synthetic mesh.  Default workload = BASELINE.json configs[2]:
the 1 M-vertex (100^3) tet block dropped on the ground plane, 10 substeps x 10 iterations.

  value : V * substeps * steps / device time, state resident in HBM (CUDA events on the
          solver's stream around the K graph launches; max over ranks)
  e2e   : the same frames through the C ABI with HOST buffers every step: sb_set_state from
          pinned host memory (H2D), sb_step, sb_read_positions + sb_read_surface (D2H)
  roofline : dominant kernel (first tile pass) timed alone with CUDA events, algorithmic
          bytes per launch / time against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (oracle/, an XPBD restatement -- the reference C# solver is
          not in the mount) on this box's host cores, bounded sample

`--impl reference` times that CPU oracle alone (all host threads) on the same workload.
N > 1 (torchrun): every rank steps its own body of the same size (independent bodies per
GPU, no communication, BASELINE.json configs[3] sharding) -> weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vertex-substeps/sec"
B_PREDICT, B_FINISH = 64.0, 64.0  # bytes per vertex


def workload(args, rank=0, world=1):
    from softbodyunity_b200 import meshgen
    if args.workload == "block":
        n = args.n
        # lowest layer 2 mm above the ground: contact within the warm-up frames (the timed frames collide)
        pos, tets, tris = meshgen.block(n, n, n, spacing=0.01, origin=(0.0, 0.002, 0.0), jitter=0.1, seed=1234 + rank)
        name = f"block{n}^3 ({n ** 3} verts) tet mesh, ground plane, S={args.substeps} I={args.iterations}"
    elif args.workload == "sphere":
        pos, tets, tris = meshgen.sphere(args.n, spacing=0.01, seed=1234 + rank)
        name = f"sphere n={args.n} ({len(pos)} verts) tet mesh, S={args.substeps} I={args.iterations}"
    elif args.workload == "bodies":
        # BASELINE.json configs[3]: the bodies are dealt to the ranks (contiguous balanced ranges)
        from softbodyunity_b200.shard import shard_range
        lo, hi = shard_range(args.n, rank, world)
        pos, tets, tris = meshgen.bodies(hi - lo, dims=(13, 13, 12), spacing=0.02, base_height=0.004, seed=1234 + rank)
        name = f"{args.n} independent 2028-vertex bodies sharded over {world} rank(s), S={args.substeps} I={args.iterations}"
    elif args.workload == "dist":
        # BASELINE.json configs[4], peer-memory version: EVERY rank plans the whole block; a rank owns a slab of it
        n = args.n
        pos, tets, tris = meshgen.block(n, n, n, spacing=0.01, origin=(0.0, 0.002, 0.0), jitter=0.1, seed=1234)
        name = (f"single block{n}^3 ({n ** 3} verts) over {world} rank(s): one address space over NVLink peer memory "
                f"(no ghosts, no exchange kernels), S={args.substeps} I={args.iterations}")
    elif args.workload == "partitioned":
        # BASELINE.json configs[4]: ONE block of n^3 vertices (default 200^3 = 8 M) cut into `world` slabs
        n = args.n
        pos, tets, tris = meshgen.block(n, n, n, spacing=0.01, origin=(0.0, 0.002, 0.0), jitter=0.1, seed=1234)
        name = (f"single block{n}^3 ({n ** 3} verts) partitioned over {world} rank(s), halo exchange over NVLink peer "
                f"memory twice per sweep, S={args.substeps} I={args.iterations}")
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    return pos, tets, tris, name


def bytes_per_substep(n_verts, n_edges, n_tets, iterations):
    """Algorithmic bytes per vertex-substep: predict 64 + I*(12 E/V + 20 T/V + 32) + finish 64."""
    b_iter = 12.0 * n_edges / n_verts + 20.0 * n_tets / n_verts + 32.0
    return B_PREDICT + iterations * b_iter + B_FINISH


def measured_traffic():
    """DRAM bytes per tile-pass launch (mean of the captured launches) from the committed ncu capture (profiles/), or None."""
    try:
        files = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
        ls = json.load(open(os.path.join(ROOT, "profiles", files[-1])))["launches"]
        return sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in ls) / len(ls), files[-1]
    except Exception:
        return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons, sampled every 100 ms from before the warm-up;
    only samples whose timestamp falls inside the marked (GPU-busy) window are kept."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or 1e18)):
                    continue
                sm.append(float(f[1]))
                mx = float(f[2])
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


def cpu_oracle_rate(pos, tets, order, off, substeps, iterations, threads, reps, roles=None):
    """vertex-substeps/s of the CPU oracle on `reps` steps of `substeps` substeps each."""
    from oracle import xpbd_oracle as orc
    m = orc.Model(pos, tets, roles=roles)
    p = orc.params(dt=(1.0 / 60.0) * substeps / 10.0, substeps=substeps, iterations=iterations)
    m.simulate(p, n_frames=1, order=order, batch_off=off, threads=threads)  # touch memory
    t0 = time.perf_counter()
    m.simulate(p, n_frames=reps, order=order, batch_off=off, threads=threads)
    dt = time.perf_counter() - t0
    return len(pos) * substeps * reps / dt, dt


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path on the host cores.  The C# solver is
    not in the mount and nothing compiles from /root/reference, so this is the oracle port."""
    if rank != 0:
        return
    from softbodyunity_b200 import SoftBody
    pos, tets, tris, name = workload(args)
    plan = SoftBody(pos, tets, tris, host_only=True)  # only the Gauss-Seidel order (colour schedule) is taken from it
    order, off = plan.schedule()
    info = plan.info()
    threads = os.cpu_count() or 1
    from oracle import xpbd_oracle as orc
    m = orc.Model(pos, tets, roles=plan.tet_roles())
    sub = 1  # one step of this arm = ONE substep (iterations sweeps) of the workload: bounded sample
    p = orc.params(dt=(1.0 / 60.0) / args.substeps, substeps=sub, iterations=args.iterations)
    for _ in range(args.warmup):
        m.simulate(p, n_frames=1, order=order, batch_off=off, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.simulate(p, n_frames=1, order=order, batch_off=off, threads=threads)
    dt = time.perf_counter() - t0
    val = len(pos) * sub * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "vertex-substeps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "step": f"1 substep x {args.iterations} iterations per step (bounded sample of the frame)",
                   "n_verts": info["n_verts"], "n_edges": info["n_edges"], "n_tets": info["n_tets"]},
        "cpu_baseline": {"value": val, "unit": "vertex-substeps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of 1 substep x {args.iterations} iterations, OpenMP over colour batches; "
                                   "CPU oracle (C), not the reference C# solver (not in the mount)"},
        "e2e": {"value": val, "unit": "vertex-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="block", choices=["block", "sphere", "bodies", "partitioned", "dist"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=100)
    ap.add_argument("--substeps", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--fast-math", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--dag", action="store_true", help="persistent tile-DAG kernel (one launch per substep)")
    ap.add_argument("--tile-cap", type=int, default=0)
    ap.add_argument("--later-tile-cap", type=int, default=0)
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--round-width", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-breakdown", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()  # nvidia-smi takes a moment to come up; samples are windowed by timestamp
    from softbodyunity_b200 import FLAG_FAST_MATH, SoftBody
    pos, tets, tris, name = workload(args, rank, world)
    flags = (FLAG_FAST_MATH if args.fast_math else 0) | (16 if args.no_pdl else 0) | (32 if args.dag else 0)
    kw = dict(substeps=args.substeps, iterations=args.iterations, flags=flags, tile_cap=args.tile_cap,
              later_tile_cap=args.later_tile_cap, block_threads=args.block_threads, round_width=args.round_width)
    part_mesh = None
    if args.workload == "partitioned" and world > 1:
        from softbodyunity_b200.partition import PartitionedBody, connect_peers, slab_partition
        (part_mesh,) = slab_partition(pos, tets, tris, world, only_rank=rank)
        V_global, T_global = len(pos), len(tets)
        del pos, tets, tris
        body = PartitionedBody(part_mesh, device=local, **kw)
        connect_peers(body, local)     # CUDA IPC handles travel over torch.distributed once; the data path is P2P stores
        sb = body.sb
    elif args.workload == "dist" and world > 1:
        from softbodyunity_b200.dist import DistBody
        V_global, T_global = len(pos), len(tets)
        body = DistBody(pos, tets, tris, device=local, **kw)
        sb = body.sb
        part_mesh = body
    else:
        sb = SoftBody(pos, tets, tris, device=local, **kw)
    info = sb.info()
    V, E, T = info["n_verts"] - info["n_ghost_verts"], info["n_edges"], info["n_tets"]
    ns = info["n_surface_verts"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ---------------------------------------------------
    clocks.mark_begin()
    sb.step(frames=args.warmup)
    sb.synchronize()
    barrier()
    ms = sb.time_frames(args.steps)
    barrier()
    ms = max_over_ranks(ms)
    # keep the GPU under the same load until the sampler has seen it (untimed)
    if part_mesh is not None:
        # ranks that share one mesh must issue the SAME number of frames: a count from the reduced time, not a clock
        sb.step(frames=max(2, int(600.0 / max(ms / args.steps, 1e-3))))
        sb.synchronize()
    else:
        t_soak = time.time()
        while time.time() - t_soak < 0.6:
            sb.step(frames=2)
            sb.synchronize()
    clocks.mark_end()
    clk = clocks.stop() if rank == 0 else None
    V_all = V if world == 1 else int(round(sum_over_ranks(V)))
    value = V_all * args.substeps * args.steps / (ms * 1e-3)

    if part_mesh is not None:
        # strong scaling of ONE mesh: report the device-timed figure only (the host-buffer protocol of the
        # single-body e2e leg would have to re-send ghosts as well; not measured for this workload)
        if sb.halo_error() or sb.dist_error():
            raise SystemExit("bench.py: a wait for a peer GPU timed out")
        if rank == 0:
            print(json.dumps({
                "metric": METRIC, "value": V_global * args.substeps * args.steps / (ms * 1e-3), "unit": "vertex-substeps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": name, "n_verts": V_global, "n_tets": T_global,
                           "own_verts_rank0": int(part_mesh.owned.sum()) if args.workload == "dist" else V,
                           "tiles_rank0": part_mesh.tiles if args.workload == "dist" else None,
                           "ghost_verts_rank0": info["n_ghost_verts"], "constraints_cut_rank0": info["constraints_cut"],
                           "math": "fast" if args.fast_math else "exact (bit-identical to CPU oracle)",
                           "tile_passes": info["n_tile_passes"], "build_seconds": info["build_seconds"]},
                "e2e": None, "gpu_launches": info["launches_per_frame"] * args.steps, "clocks": clk}), flush=True)
        dist.destroy_process_group()
        return

    # ---- end to end through the C ABI with host buffers ---------------------------------
    x4 = torch.empty((V, 4), dtype=torch.float32).pin_memory()
    v4 = torch.empty((V, 4), dtype=torch.float32).pin_memory()
    out_pos = torch.empty((V, 3), dtype=torch.float32).pin_memory()
    out_sp = torch.empty((max(ns, 1), 3), dtype=torch.float32).pin_memory()
    out_sn = torch.empty((max(ns, 1), 3), dtype=torch.float32).pin_memory()
    sb.get_state(x4, v4)

    def e2e_step():
        sb.set_state(x4, v4)            # H2D: this step's input state from pinned host memory
        sb.step()
        sb.positions(out_pos)           # D2H: all positions (mesh write-back)
        if ns:
            sb.read_surface(out_sp, out_sn)  # D2H: surface positions + normals
        sb.get_state(x4, v4)            # D2H: state handed back to the host for the next call

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = 2 * V * 16
    d2h = V * 12 + 2 * ns * 12 + 2 * V * 16
    e2e_val = V_all * args.substeps * args.steps / e2e_s

    # ---- roofline of the dominant kernel (first tile pass), timed alone --------------------
    hbm, peak_src = peaks()
    B_sub = bytes_per_substep(V, E, T, args.iterations)
    roof = None
    breakdown = None
    if info["n_tile_passes"] > 0:
        # dominant kernel = the tile pass (k_tile_rounds; n_pass launches per sweep).  Average launch duration
        # measured live: the device-timed step minus the per-vertex kernels, divided by the pass launches in it.
        n_pass = info["n_tile_passes"]
        # algorithmic bytes of one sweep (12 E + 20 T + 32 V), shared out over its n_pass launches;
        # the positions a pass re-reads from L2 are this design's overhead, not algorithmic traffic
        launch_bytes = (12.0 * E + 20.0 * T + 32.0 * V) / n_pass
        t_pred, t_fin, t_nrm = sb.time_kernel(0, 30), sb.time_kernel(1, 30), sb.time_kernel(2, 30)
        n_launch = args.substeps * args.iterations * n_pass
        step_ms = ms / args.steps
        k_ms = (step_ms - args.substeps * (t_pred + t_fin) - t_nrm) / n_launch
        alone_ms = sb.time_kernel(16, reps=30)
        ach = launch_bytes / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic()
        roof = {"bound": "hbm", "kernel": "k_tile_rounds (%d tile passes per sweep)" % n_pass, "achieved": ach, "peak": hbm, "unit": "GB/s",
                "frac": ach / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_launch": launch_bytes,
                "launch_ms": k_ms, "launch_ms_source": "(device-timed step - per-vertex kernels) / tile-pass launches in the step",
                "launches_per_step": n_launch, "share_of_step": k_ms * n_launch / step_ms,
                "pass0_alone_ms": alone_ms,
                "step_achieved": V * args.substeps * args.steps / (ms * 1e-3) * B_sub / 1e9,
                "step_frac": V * args.substeps * args.steps / (ms * 1e-3) * B_sub / 1e9 / hbm,
                "bytes_per_vertex_substep": B_sub}
        if args.kernel_breakdown:
            breakdown = {"predict_ms": sb.time_kernel(0, 30), "finish_ms": sb.time_kernel(1, 30),
                         "normals_ms": sb.time_kernel(2, 30)}
            for p in range(info["n_tile_passes"]):
                breakdown[f"pass{p}_ms"] = sb.time_kernel(16 + p, 30)
            if info["n_global_batches"]:
                breakdown["global_ms"] = sb.time_kernel(32, 30)

    # ---- CPU baseline: the oracle on the host cores, bounded sample, rank 0 ------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        order, off = sb.schedule()
        threads = os.cpu_count() or 1
        rate, secs = cpu_oracle_rate(pos, tets, order, off, 1, args.iterations, threads, reps=max(1, min(10, args.substeps)),
                                     roles=sb.tet_roles())
        cpu = {"value": rate, "unit": "vertex-substeps/s", "cores": threads, "kind": "port",
               "sample": f"{max(1, min(10, args.substeps))} substeps x {args.iterations} iterations of the same mesh and colour order "
                         f"({secs:.1f} s); CPU oracle (C, OpenMP over colour batches), not the reference C# solver (not in the mount)"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "vertex-substeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload == "bodies" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name + (" per GPU, independent bodies, no communication" if world > 1 and args.workload != "bodies" else ""),
                       "n_verts": V, "n_edges": E, "n_tets": T, "math": "fast" if args.fast_math else "exact (bit-identical to CPU oracle)",
                       "tile_passes": info["n_tile_passes"], "tiles_in_pass": info["tiles_in_pass"][:info["n_tile_passes"]],
                       "tile_cap": info["tile_cap"], "block_threads": info["block_threads"], "round_width": info["round_width"],
                       "edges_attached": info["edges_attached"], "rounds_per_sweep": sum(info["rounds_in_pass"][:info["n_tile_passes"]]),
                       "runs_per_sweep": sum(info["runs_in_pass"][:info["n_tile_passes"]]),
                       "planner": {k: os.environ[k] for k in ("SB_RECOLOUR", "SB_ATTACH_AUGMENT", "SB_MERGE_RIMS", "SB_WHOLE_BOXES", "SB_ATOM_SNAKE") if k in os.environ} or "defaults",
                       "rounds_per_tile": [round(r / max(1, t), 1) for r, t in zip(info["rounds_in_pass"][:info["n_tile_passes"]],
                                                                                  info["tiles_in_pass"][:info["n_tile_passes"]])],
                       "l2": "no flush: per-step working set (constraint streams + state) %.0f MB exceeds the 126 MB L2" %
                             ((8.0 * E + 16.0 * T + 48.0 * V) / 1e6),
                       "build_seconds": info["build_seconds"]},
            "e2e": {"value": e2e_val, "unit": "vertex-substeps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": info["launches_per_frame"] * args.steps,
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        if breakdown:
            out["kernel_breakdown_ms"] = breakdown
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
