#!/usr/bin/env python
"""bench.py -- vertex-substeps/sec of the soft-body substep path (BASELINE.json metric).

A "step" is one frame of the workload: `substeps` substeps, each with `iterations`
projection sweeps, over the synthetic mesh.  Default workload = BASELINE.json configs[2]:
the 1 M-vertex (100^3) tet block dropped on the ground plane, 10 substeps x 10 iterations.

  value : V * substeps * steps / device time, state resident in HBM (CUDA events on the
          solver's stream around the K graph launches; max over ranks)
  e2e   : the same frames through the C ABI with HOST buffers every step: sb_write_packed (x4 | v4 from
          pinned host memory, H2D), sb_step, sb_read_packed (x4 | v4 | surface positions | normals, D2H)
  roofline : the tile launches (they carry the whole step: predict / finish run inside them): algorithmic
          bytes per launch / mean launch time inside the device-timed step, against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (oracle/, an XPBD restatement -- the reference C# solver is
          not in the mount) on this box's host cores, bounded sample

`--impl reference` times that CPU oracle alone (all host threads) on the same workload.
N > 1 (torchrun): the path that shards -- ONE 200^3 = 8 M-vertex block spread over all the GPUs through
NVLink peer memory (BASELINE.json configs[4], strong scaling; `--workload dist`), with the 4096-body batch
of configs[3] sharded over the ranks as a secondary figure in the same line (`bodies`).
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


def measured_traffic(info=None):
    """DRAM bytes per tile-pass launch (mean of the captured launches) from the newest committed ncu capture
    (profiles/*_traffic.json), or (None, reason): a capture of another plan (rounds per sweep / tiles per pass differ
    from what runs now) is refused rather than quoted."""
    try:
        files = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
        doc = json.load(open(os.path.join(ROOT, "profiles", files[-1])))
        ls = doc["launches"]
        if info is not None:
            np_ = info["n_tile_passes"]
            want = {"rounds_per_sweep": int(sum(info["rounds_in_pass"][:np_])), "tiles_in_pass": [int(t) for t in info["tiles_in_pass"][:np_]]}
            have = {k: doc.get(k) for k in want}
            if have != want:
                return None, f"{files[-1]} is of another plan ({have} captured, {want} running): not quoted"
        return sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in ls) / len(ls), files[-1]
    except Exception as e:
        return None, f"no usable capture under profiles/ ({type(e).__name__})"


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


def solver_flags(args):
    from softbodyunity_b200 import FLAG_DAG, FLAG_FAST_MATH, FLAG_NO_FUSE, FLAG_NO_PDL, FLAG_NO_SNAKE
    return ((FLAG_FAST_MATH if args.fast_math else 0) | (FLAG_NO_PDL if args.no_pdl else 0) | (FLAG_DAG if args.dag else 0) |
            (FLAG_NO_SNAKE if args.no_snake else 0) | (FLAG_NO_FUSE if args.no_fuse else 0))


def plan_options(args):
    """Planner options of a run (the same for the native arm and for the plan the reference arm takes its order from)."""
    kw = dict(tile_cap=args.tile_cap, later_tile_cap=args.later_tile_cap, block_threads=args.block_threads, round_width=args.round_width)
    if args.workload == "dist":
        # ONE plan for every N: the box grid and the block numbering are those of an 8-way split (2 and 4 ranks take
        # unions of its blocks), the CTA width is pinned -- so the Gauss-Seidel order, and with it the state checksum,
        # is the same on 1, 2, 4 and 8 GPUs
        kw["dist_ranks"] = 0 if args.slabs else 8
        kw["block_threads"] = args.block_threads or 160
    return kw


def describe_config(args, info, name, n_verts, world, replicas=False):
    """`config` of the JSON line: the workload and the plan it runs on.  Both arms print the same object for the same
    arguments (the reference arm replays the plan's Gauss-Seidel order on the host cores)."""
    npass = info["n_tile_passes"]
    ws = (8.0 * info["n_edges"] + 16.0 * info["n_tets"] + 48.0 * n_verts) / 1e6  # MB, per GPU-set
    cfg = {"workload": name + (" per GPU, independent bodies, no communication" if replicas and world > 1 else ""),
           "n_verts": n_verts, "n_edges": info["n_edges"], "n_tets": info["n_tets"],
           "math": "fast" if args.fast_math else "exact (bit-identical to CPU oracle)",
           "tile_passes": npass, "tiles_in_pass": info["tiles_in_pass"][:npass],
           "tile_cap": info["tile_cap"], "block_threads": info["block_threads"], "round_width": info["round_width"],
           "edges_attached": info["edges_attached"], "rounds_per_sweep": sum(info["rounds_in_pass"][:npass]),
           "runs_per_sweep": sum(info["runs_in_pass"][:npass]),
           "order": "same pass order every iteration" if args.no_snake else "snake: odd iterations run the tile passes backwards",
           "launches": "one per pass occurrence, separate predict / finish" if args.no_fuse else
                       "consecutive occurrences of a pass fused; predict / finish inside the tile launches at substep boundaries",
           "planner": {k: os.environ[k] for k in ("SB_RECOLOUR", "SB_ATTACH_AUGMENT", "SB_MERGE_RIMS", "SB_WHOLE_BOXES", "SB_ATOM_SNAKE") if k in os.environ} or "defaults",
           "rounds_per_tile": [round(r / max(1, t), 1) for r, t in zip(info["rounds_in_pass"][:npass], info["tiles_in_pass"][:npass])],
           "l2": ("no flush: per-step working set (constraint streams + state) %.0f MB exceeds the 126 MB L2" if ws > 126.0 else
                  "NOT A BENCH CONFIGURATION: per-step working set %.0f MB fits in the 126 MB L2 and nothing flushes it") % ws}
    if world > 1 and args.workload in ("dist", "partitioned"):
        cfg.update(comm="NVLink peer memory (CUDA IPC): tiles read / write their vertex runs in the owner's HBM; epoch words "
                        "order the zone tiles of neighbouring ranks; no collective on the data path (NCCL: rendezvous, timing)"
                        if args.workload == "dist" else "ghost vertices + P2P halo kernels twice per sweep",
                   partition="slabs of the box order" if args.slabs else "compact blocks of boxes (recursive bisection)")
    return cfg


def cpu_oracle_rate(pos, tets, sched, substeps, iterations, threads, reps, roles=None):
    """vertex-substeps/s of the CPU oracle on `reps` steps of `substeps` substeps each."""
    from oracle import xpbd_oracle as orc
    m = orc.Model(pos, tets, roles=roles)
    p = orc.params(dt=(1.0 / 60.0) * substeps / 10.0, substeps=substeps, iterations=iterations)
    m.simulate(p, n_frames=1, threads=threads, **sched)  # touch memory
    t0 = time.perf_counter()
    m.simulate(p, n_frames=reps, threads=threads, **sched)
    dt = time.perf_counter() - t0
    return len(pos) * substeps * reps / dt, dt


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path on the host cores.  The C# solver is
    not in the mount and nothing compiles from /root/reference, so this is the oracle port."""
    if rank != 0:
        return
    from softbodyunity_b200 import SoftBody
    # (the mesh of the distributed workload does not depend on the rank count; its name does, and both arms print it)
    pos, tets, tris, name = workload(args, 0, world if args.workload in ("dist", "partitioned") else 1)
    # only the Gauss-Seidel order (colour schedule) is taken from the plan; the same plan options as the native arm
    plan = SoftBody(pos, tets, tris, host_only=True, substeps=args.substeps, iterations=args.iterations, flags=solver_flags(args),
                    **plan_options(args))
    sched = plan.schedule_kw()
    info = plan.info()
    threads = os.cpu_count() or 1
    from oracle import xpbd_oracle as orc
    m = orc.Model(pos, tets, roles=plan.tet_roles())
    sub = 1  # one step of this arm = ONE substep (iterations sweeps) of the workload: bounded sample
    p = orc.params(dt=(1.0 / 60.0) / args.substeps, substeps=sub, iterations=args.iterations)
    for _ in range(args.warmup):
        m.simulate(p, n_frames=1, threads=threads, **sched)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.simulate(p, n_frames=1, threads=threads, **sched)
    dt = time.perf_counter() - t0
    val = len(pos) * sub * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "vertex-substeps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": describe_config(args, info, name, info["n_verts"], world),
        "step_sample": f"one step of this arm = 1 substep x {args.iterations} iterations of the workload (bounded sample of the frame; "
                       "the unit, vertex-substeps/s, is the same)",
        "cpu_baseline": {"value": val, "unit": "vertex-substeps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of 1 substep x {args.iterations} iterations, OpenMP over colour batches; "
                                   "CPU oracle (C), not the reference C# solver (not in the mount)"},
        "e2e": {"value": val, "unit": "vertex-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "plan_build_seconds": info["build_seconds"],
    }
    print(json.dumps(out), flush=True)


def checksum_key(cfg, args):
    """What a state checksum is a function of: mesh, stepping and the plan (the Gauss-Seidel order)."""
    return "%s n=%d S=%d I=%d | tiles %s | rounds %d | bt %d | %s" % (
        args.workload, args.n, args.substeps, args.iterations, cfg["tiles_in_pass"], cfg["rounds_per_sweep"], cfg["block_threads"],
        "fast" if args.fast_math else "exact")


def oracle_checksum(cfg, args, frames):
    """The CPU oracle's checksum of this workload after `frames` frames from the committed fixture
    (tests/golden/dist_checksum.json, made by tests/golden/make_dist_checksum.py), or None if it holds none."""
    try:
        doc = json.load(open(os.path.join(ROOT, "tests", "golden", "dist_checksum.json")))
        return doc[checksum_key(cfg, args)]["after_frames"].get(str(frames))
    except Exception:
        return None


def state_checksum(x4_owned, world, dist, torch):
    """Order-independent checksum of the positions (sum of the 32-bit words of x4, high and low halves apart),
    over all ranks: the same mesh stepped the same number of frames gives the same value on any number of GPUs."""
    w = np.ascontiguousarray(x4_owned, np.float32).view(np.uint32).astype(np.int64)
    t = torch.tensor([int((w >> 16).sum()), int((w & 0xffff).sum())], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return "%x-%x" % (int(t[0].item()), int(t[1].item()))


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None, choices=["block", "sphere", "bodies", "partitioned", "dist", "replicas"],
                    help="default: block (1 M vertices) on one GPU; on N > 1 GPUs `dist`, ONE 200^3 = 8 M-vertex block over all of them")
    ap.add_argument("--n", "--size", dest="n", type=int, default=None)
    ap.add_argument("--substeps", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--fast-math", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--dag", action="store_true", help="persistent tile-DAG kernel (one launch per substep)")
    ap.add_argument("--no-snake", action="store_true", help="A/B: same pass order in every iteration")
    ap.add_argument("--no-fuse", action="store_true", help="A/B: one launch per tile pass, separate predict / finish kernels")
    ap.add_argument("--slabs", action="store_true", help="A/B (dist): ranks as slabs of the default box order, not compact blocks")
    ap.add_argument("--tile-cap", type=int, default=0)
    ap.add_argument("--later-tile-cap", type=int, default=0)
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--round-width", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bodies", action="store_true", help="skip the secondary 4096-body measurement (BASELINE.json configs[3])")
    ap.add_argument("--bodies-n", type=int, default=4096)
    ap.add_argument("--kernel-breakdown", action="store_true")
    args = ap.parse_args(argv)
    args.warmup = max(args.warmup, 3)
    return args


def main(argv=None):
    args = parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        # N > 1: the path that SHARDS -- one mesh spread over the GPUs (BASELINE.json configs[4]), strong scaling
        args.workload = "block" if world == 1 else "dist"
    if args.workload == "replicas":  # (round 1's N > 1 default: one independent 1 M body per GPU)
        args.workload, replicas = "block", True
    else:
        replicas = False
    if args.n is None:
        args.n = {"block": 100, "sphere": 58, "bodies": args.bodies_n, "dist": 200, "partitioned": 200}[args.workload]

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
    from softbodyunity_b200 import SoftBody
    pos, tets, tris, name = workload(args, rank, world)
    flags = solver_flags(args)
    kw = dict(substeps=args.substeps, iterations=args.iterations, flags=flags, **plan_options(args))
    shared = None      # a body shared by all ranks (one mesh over the GPUs)
    if args.workload == "partitioned" and world > 1:
        from softbodyunity_b200.partition import PartitionedBody, connect_peers, slab_partition
        (part,) = slab_partition(pos, tets, tris, world, only_rank=rank)
        V_global, T_global = len(pos), len(tets)
        del pos, tets, tris
        shared = PartitionedBody(part, device=local, **kw)
        connect_peers(shared, local)     # CUDA IPC handles travel over torch.distributed once; the data path is P2P stores
        sb = shared.sb
    elif args.workload == "dist":
        # (plan_options: ONE plan for every N)
        if world > 1:
            from softbodyunity_b200.dist import DistBody
            V_global, T_global = len(pos), len(tets)
            shared = DistBody(pos, tets, tris, device=local, **kw)
            sb = shared.sb
        else:
            sb = SoftBody(pos, tets, tris, device=local, **kw)
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

    def timed_frames(body, steps, warmup):
        """device time (ms, max over ranks) of `steps` frames after `warmup` frames, barrier + synchronize both sides"""
        body.step(frames=warmup)
        body.synchronize()
        barrier()
        t = body.time_frames(steps)
        barrier()
        return max_over_ranks(t)

    # ---- device-resident throughput ---------------------------------------------------
    clocks.mark_begin()
    ms = timed_frames(sb, args.steps, args.warmup)
    checksum = None
    if shared is not None and args.workload == "dist":
        x4o = sb.unpack_frame(sb.read_packed())[0]
        checksum = state_checksum(x4o, world, dist, torch)
    elif world == 1 and args.workload in ("dist", "block", "sphere"):
        # one GPU: for `dist` the value every N must reproduce; for the headline mesh the value the CPU oracle gives
        checksum = state_checksum(sb.get_state()[0], 1, dist, torch)
    # keep the GPU under the same load until the sampler has seen it (untimed)
    if shared is not None:
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
    if shared is not None:
        V_all = V_global
        if sb.halo_error() or sb.dist_error():
            raise SystemExit("bench.py: a wait for a peer GPU timed out")
    else:
        V_all = V if world == 1 else int(round(sum_over_ranks(V)))
    value = V_all * args.substeps * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers, one packed copy each way -------------------
    # every step: H2D of the step's input state (x4 + v4 of the vertices this rank owns) from pinned memory, sb_step,
    # D2H of the frame (state + surface positions + surface normals) into pinned memory; sb_read_packed synchronises
    e2e = None
    if args.workload != "partitioned" or shared is None:
        n_own, ns_own, b_in, b_out = sb.packed_sizes()
        host = torch.empty(b_out, dtype=torch.uint8).pin_memory()
        sb.read_packed(host)

        def e2e_step():
            sb.write_packed(host)   # H2D (the first b_in bytes: x4 | v4)
            sb.step()
            sb.read_packed(host)    # D2H: x4 | v4 | surface xyz | surface normals

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": V_all * args.substeps * args.steps / e2e_s, "unit": "vertex-substeps/s",
               "h2d_bytes_per_step": int(sum_over_ranks(b_in)), "d2h_bytes_per_step": int(sum_over_ranks(b_out)),
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "protocol": "sb_write_packed (x4 | v4 from pinned host memory) -> sb_step -> sb_read_packed (x4 | v4 | surface xyz | "
                           "surface normals into pinned host memory), one copy and one synchronisation each way"
                           + (", every rank its own vertices" if shared is not None else "")}
        if shared is not None and sb.dist_error():
            raise SystemExit("bench.py: a wait for a peer GPU timed out")

    # ---- supplementary: the frame as the managed component drives it (not the headline e2e above) -------------------
    # The drop-in boundary per frame is sb_step(dt) and a read-back of the surface mesh (positions + normals of the
    # surface vertices into pinned host memory); the state stays on the device.  Reported beside `e2e`, never instead.
    e2e_component = None
    if shared is None and ns > 0:
        try:
            sp = torch.empty((ns, 3), dtype=torch.float32).pin_memory()
            sn = torch.empty((ns, 3), dtype=torch.float32).pin_memory()

            def component_step():
                sb.step(1.0 / 60.0)       # FixedUpdate: dt is the frame's only input
                sb.read_surface(sp, sn)   # LateUpdate: D2H of the mesh the renderer needs; synchronises

            for _ in range(2):
                component_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                component_step()
            barrier()
            comp_s = max_over_ranks(time.perf_counter() - t0)
            e2e_component = {"value": V_all * args.substeps * args.steps / comp_s, "unit": "vertex-substeps/s", "ms_per_step": 1e3 * comp_s / args.steps,
                             "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(sum_over_ranks(24 * ns)),
                             "protocol": "sb_step(dt) -> sb_read_surface (surface positions + normals into pinned host memory, one synchronisation): "
                                         "what the C# component moves per frame; state resident on the device.  Supplementary: the headline "
                                         "end-to-end figure is `e2e` (full state in and out every frame)"}
        except Exception as exc:  # a supplementary figure must not cost the line
            e2e_component = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    hbm, peak_src = peaks()
    B_sub = bytes_per_substep(V_all if shared is not None else V, E, T, args.iterations)
    roof = None
    breakdown = None
    if shared is not None:
        # whole-step figure only (sb_time_kernel would run launches the peers do not run)
        ach = value * B_sub / 1e9
        roof = {"bound": "hbm", "kernel": "k_tile_rounds (whole step, all ranks)", "achieved": ach, "peak": hbm * world, "unit": "GB/s",
                "frac": ach / (hbm * world), "traffic": None, "traffic_source": "not captured for the multi-GPU run (ncu is a one-GPU tool here)",
                "peak_source": peak_src + " x n_gpus", "bytes_per_vertex_substep": B_sub,
                "launch_ms": ms / args.steps / max(1, info["launches_per_frame"]),
                "launch_ms_source": "device-timed step / launches in it (max over ranks)"}
    elif info["n_tile_passes"] > 0:
        n_pass = info["n_tile_passes"]
        # The frame program (sb_frame_program): consecutive occurrences of a pass share one launch, and predict /
        # collide + velocity update run inside the tile launches at the substep boundaries, so the tile launches
        # carry the whole algorithmic traffic of the step (S * (64 + I * (12 E + 20 T + 32 V) / V + 64)
        # per vertex) minus what separate per-vertex kernels, if any, carry (64 B per vertex each).
        prog = sb.frame_program()
        n_launch = int((prog[:, 0] == 2).sum())
        n_pred, n_fin, n_nrm = int((prog[:, 0] == 0).sum()), int((prog[:, 0] == 1).sum()), int((prog[:, 0] == 6).sum())
        t_pred, t_fin, t_nrm = sb.time_kernel(0, 30), sb.time_kernel(1, 30), sb.time_kernel(2, 30)
        step_ms = ms / args.steps
        k_ms = (step_ms - n_pred * t_pred - n_fin * t_fin - n_nrm * t_nrm) / n_launch
        launch_bytes = (B_sub * V * args.substeps - 64.0 * V * (n_pred + n_fin)) / n_launch
        alone_ms = sb.time_kernel(16, reps=30)
        ach = launch_bytes / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(info)
        roof = {"bound": "hbm", "kernel": "k_tile_rounds (%d tile passes per sweep, %d launches per step)" % (n_pass, n_launch),
                "achieved": ach, "peak": hbm, "unit": "GB/s",
                "frac": ach / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_launch": launch_bytes,
                "launch_ms": k_ms, "launch_ms_source": "(device-timed step - separate per-vertex kernels) / tile-pass launches in the step; "
                                                       "bytes_per_launch = algorithmic bytes of the step the tile launches carry / their number",
                "launches_per_step": n_launch, "separate_vertex_kernels_per_step": n_pred + n_fin,
                "share_of_step": k_ms * n_launch / step_ms,
                "pass0_alone_ms": alone_ms,
                "step_achieved": V * args.substeps * args.steps / (ms * 1e-3) * B_sub / 1e9,
                "step_frac": V * args.substeps * args.steps / (ms * 1e-3) * B_sub / 1e9 / hbm,
                "step_frac_of_nominal_8000": V * args.substeps * args.steps / (ms * 1e-3) * B_sub / 1e9 / 8000.0,
                "bytes_per_vertex_substep": B_sub}
        if args.kernel_breakdown:
            breakdown = {"predict_ms": t_pred, "finish_ms": t_fin, "normals_ms": t_nrm}
            for p in range(info["n_tile_passes"]):
                breakdown[f"pass{p}_ms"] = sb.time_kernel(16 + p, 30)
            if info["n_global_batches"]:
                breakdown["global_ms"] = sb.time_kernel(32, 30)

    # ---- CPU baseline: the oracle on the host cores, bounded sample, rank 0 at N = 1 ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, secs = cpu_oracle_rate(pos, tets, sb.schedule_kw(), 1, args.iterations, threads, reps=max(1, min(10, args.substeps)),
                                     roles=sb.tet_roles())
        cpu = {"value": rate, "unit": "vertex-substeps/s", "cores": threads, "kind": "port",
               "sample": f"{max(1, min(10, args.substeps))} substeps x {args.iterations} iterations of the same mesh and colour order "
                         f"({secs:.1f} s); CPU oracle (C, OpenMP over colour batches), not the reference C# solver (not in the mount)"}

    own_rank0 = int(shared.owned.sum()) if (shared is not None and args.workload == "dist") else V
    tiles_rank0 = shared.tiles if (shared is not None and args.workload == "dist") else None
    launches = info["launches_per_frame"]
    del sb, shared
    import gc
    gc.collect()

    # ---- secondary: BASELINE.json configs[3], 4096 independent 2 k-vertex bodies sharded over the ranks --------
    bodies = None
    if not args.no_bodies and args.workload in ("block", "dist"):
        from softbodyunity_b200 import meshgen
        from softbodyunity_b200.shard import shard_range
        lo, hi = shard_range(args.bodies_n, rank, world)
        bp, bt_, bf = meshgen.bodies(hi - lo, dims=(13, 13, 12), spacing=0.02, base_height=0.004, seed=1234 + rank)
        bsb = SoftBody(bp, bt_, bf, device=local, substeps=args.substeps, iterations=args.iterations, flags=flags & ~32)
        bsteps = max(2, args.steps // 4)
        bms = timed_frames(bsb, bsteps, 3)
        bi = bsb.info()
        bV = int(round(sum_over_ranks(bi["n_verts"])))
        bval = bV * args.substeps * bsteps / (bms * 1e-3)
        bB = bytes_per_substep(bi["n_verts"], bi["n_edges"], bi["n_tets"], args.iterations)
        bodies = {"workload": f"{args.bodies_n} independent 2028-vertex bodies sharded over {world} rank(s), no communication "
                              f"(BASELINE.json configs[3]), S={args.substeps} I={args.iterations}",
                  "value": bval, "unit": "vertex-substeps/s", "scaling": "strong", "n_verts": bV, "steps": bsteps, "ms_per_step": bms / bsteps,
                  "launches_per_step": bi["launches_per_frame"], "bodies_rank0": hi - lo,
                  "roofline_frac": bval * bB / 1e9 / (hbm * world), "bytes_per_vertex_substep": bB, "build_seconds": bi["build_seconds"]}
        del bsb

    if rank == 0:
        cfg = describe_config(args, info, name, V_all if world > 1 and not replicas else V, world, replicas)
        out = {
            "metric": METRIC, "value": value, "unit": "vertex-substeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if (replicas and world > 1) else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": e2e, "gpu_launches": launches * args.steps,
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        out["plan_build_seconds"] = info["build_seconds"]
        if e2e_component is not None:
            out["e2e_component"] = e2e_component
        if world > 1 and args.workload in ("dist", "partitioned"):
            out["distribution"] = {"own_verts_rank0": own_rank0, "tiles_rank0": tiles_rank0}
        if checksum is not None:
            out["state_checksum"] = {"after_frames": args.warmup + args.steps, "x4_words_hi_lo": checksum,
                                     "meaning": "sum of the 32-bit words of every vertex's (x, y, z, 1/m), high and low halves; the same for any "
                                                "number of GPUs (the execution order is the plan's), and for `--workload dist` on one GPU with the same plan"}
            want = oracle_checksum(cfg, args, args.warmup + args.steps)
            out["state_checksum"]["cpu_oracle"] = want  # None: the fixture holds no run of this workload / frame count
            out["state_checksum"]["matches_cpu_oracle"] = None if want is None else (want == checksum)
        if bodies is not None:
            out["bodies"] = bodies
        if breakdown:
            out["kernel_breakdown_ms"] = breakdown
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
