#!/usr/bin/env python
"""bench.py -- agent-replans/s of the replan hot path (LSC + SFC + goal + QP for every agent) on the
synthetic 4096-agent 3-D forest swarm of BASELINE.json (configs[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--agents 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one replan of all agents.  The swarm is rolled out for `--settle` untimed steps first so that the
timed steps are in-transit replans (non-trivial LSC / SFC / QP active sets), and the per-step waypoints of
the whole run are recorded by an untimed pilot rollout (the waypoint provider is host-side and out of the
hot path's scope).  The path is deterministic, so the timed replays see exactly the pilot's states.

  value : agents * K / t, inputs resident in HBM (device-chained plan -> advance [-> record exchange over NVLink peer memory,
          or NCCL all-gather with --exchange nccl]),
          timed with CUDA events on the launching stream, max over ranks.
  e2e   : same metric through the host-facing C-ABI calls with HOST buffers: every step uploads pos / vel /
          acc / waypoint of every agent from pinned memory (dlsc_set_agents), runs dlsc_step and reads every
          trajectory back (dlsc_get_traj).
  scaling: strong (the 4096-agent swarm is sharded over the ranks; one all-gather of the agent records
          per step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from dlsc_gc_planner_b200 import capi, missions, sharding  # noqa: E402

METRIC = "agent-replans/sec (LSC+SFC+QP), synthetic 4096-agent 3D forest"
UNIT = "agent-replans/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=4096)
    ap.add_argument("--half-extent", type=float, default=None, help="world half size in m (default: 1 agent/m^2)")
    ap.add_argument("--max-nbr", type=int, default=96)
    ap.add_argument("--settle", type=int, default=25, help="untimed rollout steps before the timed region")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mc-missions", type=int, default=128, help="Monte-Carlo leg: independent empty50 missions per GPU (0: skip)")
    ap.add_argument("--mc-steps", type=int, default=20)
    ap.add_argument("--closed-loop-steps", type=int, default=4, help="steps of the closed loop with the waypoint provider (0: skip)")
    ap.add_argument("--dyn-obstacles", type=int, default=8, help="dynamic-obstacle leg (1 GPU): obstacles flying through the swarm (0: skip)")
    ap.add_argument("--dyn-steps", type=int, default=20)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU record exchange: stores over NVLink peer memory (default) or NCCL all-gather")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
def workload_config(args, cfg, m):
    """The `config` object of the JSON line: a function of the command line only, so that both arms (--impl ours /
    --impl reference) print the identical dict for the same flags."""
    inv = 1.0 / cfg.world_res
    dims = [int(np.floor(inv * float(np.float32(m.world_max[k])))) - int(np.floor(inv * float(np.float32(m.world_min[k])))) + 1
            for k in range(3)]
    ncell = dims[0] * dims[1] * dims[2]
    return {"workload": "synthetic %d-agent 3D random-forest swarm (BASELINE configs[3]): M=%d n=%d dim=%d, SFC on a %dx%dx%d "
                        "EDT grid, comm range %g, max_nbr %d" % (m.n_agents, cfg.M, cfg.n, cfg.dim, dims[0], dims[1], dims[2],
                                                                  cfg.comm_range, args.max_nbr),
            "agents": int(m.n_agents), "gpus": int(args.gpus), "agents_per_gpu": int(-(-m.n_agents // max(args.gpus, 1))),
            "settle_steps": int(args.settle), "seed": 4096,
            "parallelism": "agents sharded x%d, one all-gather of the agent records per step" % args.gpus,
            "l2": "per-step working set (EDT grid %.0f MB + scratch) exceeds L2; inputs change every step" % (ncell * 16 / 1e6)}


def make_world(args):
    cfg = missions.PlannerConfig.forest3d()
    h = args.half_extent if args.half_extent else 0.5 * float(np.sqrt(args.agents))
    m = missions.synthetic_forest(n_agents=args.agents, half_extent=h, seed=4096)
    return cfg, m


class Clocks:
    """SM clock / throttle-reason sampler for the timed regions (B200_PROFILING.md clocks line).  The timed regions
    last tens of milliseconds, far below nvidia-smi's start-up time, so the sampler is an in-process NVML thread
    (same counters as `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`) polling every ~2 ms;
    only samples taken inside a marked window (`begin()` .. `end()`, i.e. while the GPU is under the timed load)
    are reported.  Falls back to an nvidia-smi -lms loop when NVML cannot be opened."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        import threading
        self.samples, self.windows, self._t0 = [], [], None
        self.max_mhz, self.stop_flag, self.smi = None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception:
            self.nv = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            try:
                self.smi = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                             "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                            stderr=subprocess.DEVNULL)
            except Exception:
                self.smi = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((time.perf_counter(), mhz, rs, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def begin(self):
        self._t0 = time.perf_counter()

    def end(self):
        self.windows.append((self._t0, time.perf_counter()))

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if self.nv is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
            use = inside if inside else self.samples
            reasons = set()
            for s in use:
                for bit, name in self.REASONS.items():
                    if s[2] & bit:
                        reasons.add(name)
            if use:
                out["sm_mhz"] = statistics.median(s[1] for s in use)
                out["sm_mhz_min"] = min(s[1] for s in use)
                pw = [s[3] for s in use if s[3] is not None]
                out["power_w_max"] = max(pw) if pw else None
            out["samples"] = len(inside)
            out["samples_total"] = len(self.samples)
            out["reasons"] = sorted(reasons)
            out["how"] = "NVML polled every ~2 ms in-process; samples inside the timed windows (resident + e2e)"
            return out
        if self.smi is None:
            return out
        self.smi.terminate()
        try:
            self.smi.wait(timeout=5)
        except Exception:
            self.smi.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip() == "Active":
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        out["how"] = "nvidia-smi -lms 20 from before the warm-up to the end of the e2e region"
        return out


def pinned(shape, dtype):
    import torch
    t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
    return t, t.numpy()


# ------------------------------------------------------------------------------------------------------
def pilot_rollout(cfg, m, args, device, n_steps):
    """Untimed: roll the whole swarm out on this rank's GPU, recording per-step waypoints and host states,
    and the planner state at the start of the timed region."""
    pl = capi.SwarmPlanner(cfg, m, max_nbr=args.max_nbr, device=device)
    pl.build_edt(m.boxes)             # distance grid built on the device from the mission's obstacle boxes
    occupied = missions.occupied_nodes(m.boxes, cfg.grid_res)
    N = m.n_agents
    wp = pl.start.copy()
    goal_des = m.goal.astype(np.float32)
    rec = {"wp": [], "pos": [], "vel": [], "acc": []}
    snap = None
    traj = None
    fails, overflow = 0, 0
    for t in range(args.settle + n_steps):
        pos, vel, acc = pl.state()
        goal_cur = pl.goal()
        wp = missions.next_waypoints(wp, goal_cur, goal_des, traj, pos, cfg, occupied)
        if t == args.settle:
            snap = {"records": pl.get_records(), "sfc": pl.sfc(), "acc": acc.copy(), "seq": pl.seq}
        if t >= args.settle:
            rec["wp"].append(wp.copy()); rec["pos"].append(pos); rec["vel"].append(vel); rec["acc"].append(acc)
        pl.set_agents(waypoint=wp)
        pl.plan()
        traj = pl.traj()
        st = pl.status()
        fails += int(((st & capi.FAIL_MASK) != 0).sum())
        overflow = max(overflow, int(((st & capi.NBR_OVERFLOW) != 0).sum()))
        pl.advance()
    snap["final_traj"] = traj
    snap["nbr_overflow"] = overflow
    snap["fails"] = fails
    snap["dist_to_goal"] = float(np.mean(np.max(np.abs(pl.state()[0] - goal_des), axis=1)))
    snap["edt"] = pl.get_edt()        # the same grid arrays for the CPU baseline (oracle)
    pl.close()
    return {k: np.array(v) for k, v in rec.items()}, snap


def restore(pl, snap, sl):
    pl.set_records(0, snap["records"])
    pl.set_sfc(snap["sfc"][sl], np.zeros(pl.NL, np.uint8))
    pl.set_agents(acc=snap["acc"][sl])
    pl.seq = snap["seq"]


def qp_flops(cfg, nbr_cnt, iters, np_rows_nnz):
    """Algorithmic flops of the dual active-set QP kernel: every agent evaluates all its rows once per
    scan (iters + 1 scans): 12 flops per LSC row (b = -(d + n.anchor), v = -n.x - b), 4 per non-zero of the pattern
    rows; every iteration adds two H^-1 products over the ny x nyd block structure; plus the tabulated
    unconstrained optimum (8 flops per unknown) and the objective (12 x 6 per control-point coordinate)."""
    M, P, D = cfg.M, cfg.n + 1, cfg.dim
    nyd = 3 * M - 2
    ny = D * nyd
    r_lsc = nbr_cnt.astype(np.float64) * (M * P - 3)
    scans = iters.astype(np.float64) + 1.0
    per_scan = 12.0 * r_lsc + 4.0 * np_rows_nnz
    per_iter = 4.0 * ny * nyd
    fixed = 8.0 * ny + 12.0 * 6 * D * M * P
    return float(np.sum(scans * per_scan + iters * per_iter + fixed))


def lsc_flops(cfg, pairs, gjk_iters):
    """K [(M-1)(F_gjk + 10 P) + F_seg]; F_gjk = iterations x (2(5P+5) + 12 + S), with the measured
    iteration count and S = 25 (the sub-simplex step that dominates: 1.5 iterations per call on this workload)."""
    M, P = cfg.M, cfg.n + 1
    return float(gjk_iters * (2 * (5 * P + 5) + 12 + 25) + pairs * (M - 1) * 10 * P + pairs * 150)


def pair_nnz(cfg):
    """non-zeros of the velocity / acceleration / comm-range rows in x-space, both sides"""
    M, n, D = cfg.M, cfg.n, cfg.dim
    vel = D * (M * n - 2) * 2 * 2
    acc = D * (M * (n - 1) - 1) * 3 * 2
    comm = D * (M * (M + 1) // 2) * 2 * 2 if cfg.comm_range > 0 else 0
    return vel + acc + comm


# ------------------------------------------------------------------------------------------------------
def run_montecarlo(args, dev, world, rank, dist):
    """BASELINE configs[4]: Monte-Carlo batch of independent empty50 missions replanning in lockstep, `--mc-missions` per
    GPU (1024 missions on 8 GPUs), weak scaling, no data-path collective (replicas only).  Mission g of the
    batch is the reference's missions/empty50 #(g mod 30 + 1) with seeded goal noise (Mission::addNoise, max_noise 0.2,
    seed g).  Same protocol as the main workload: untimed pilot rollout records the waypoints, the timed pass replays them
    on the device (plan -> advance chained, CUDA events, max over ranks)."""
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", "missions_all.npz"))
    cfg = missions.PlannerConfig.empty()

    def base(idx):
        g = lambda f: z["empty50/%d/%s" % (idx, f)]
        return missions.Mission(g("world_min"), g("world_max"), g("start"), g("goal"), g("radius"), g("downwash"), g("max_vel"),
                                g("max_acc"), g("nominal_vel"), g("boxes"))
    nm = args.mc_missions
    ms = [missions.add_goal_noise(base((rank * nm + i) % 30 + 1), 0.2, cfg.dim, seed=rank * nm + i) for i in range(nm)]
    batch, group = missions.concat_missions(ms)
    n_each = ms[0].n_agents
    N, Kn = batch.n_agents, n_each - 1
    W, T, settle = 3, args.mc_steps, 10

    def planner():
        pl = capi.SwarmPlanner(cfg, batch, max_nbr=Kn, device=dev.index)
        pl.set_groups(group)
        return pl
    pl = planner()
    wp = pl.start.copy()
    goal_des = batch.goal.astype(np.float32)
    traj, rec_wp, snap, fails = None, [], None, 0
    for t in range(settle + W + T):
        pos, vel, acc = pl.state()
        goal_cur = pl.goal()
        for i in range(nm):
            q = slice(i * n_each, (i + 1) * n_each)
            wp[q] = missions.next_waypoints(wp[q], goal_cur[q], goal_des[q], None if traj is None else traj[q], pos[q], cfg)
        if t == settle:
            snap = {"records": pl.get_records(), "acc": acc.copy(), "seq": pl.seq}
        if t >= settle:
            rec_wp.append(wp.copy())
        pl.set_agents(waypoint=wp)
        pl.plan()
        traj = pl.traj()
        fails += int(((pl.status() & capi.FAIL_MASK) != 0).sum())
        pl.advance()
    final = traj
    pl.close()
    pl = planner()
    pl.set_stream(torch.cuda.current_stream().cuda_stream)
    wp_dev = torch.from_numpy(np.ascontiguousarray(np.array(rec_wp))).to(dev)

    def restore_mc():
        pl.set_records(0, snap["records"]); pl.set_agents(acc=snap["acc"]); pl.seq = snap["seq"]

    def run(n0, n):
        for t in range(n0, n0 + n):
            pl.set_waypoints_device(wp_dev[t].data_ptr())
            pl.plan(); pl.advance()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    restore_mc(); barrier()
    run(0, W); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(W, T); e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    exact = bool(np.array_equal(pl.traj(), final))
    restore_mc(); torch.cuda.synchronize()
    run(0, W); pl.enable_timing(True); run(W, T); torch.cuda.synchronize()
    stage_ms, _ = pl.timings()
    pl.enable_timing(False)
    pl.close()
    tt = torch.tensor([ms_total, float(fails), 1.0 if exact else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = tt.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        ms_total, exact = float(mx[0].item()), bool(mn[2].item() > 0.5)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        fails = int(sm[1].item())
    return {"workload": "Monte-Carlo batch (BASELINE configs[4]): %d independent empty50 missions per GPU x %d GPUs = %d missions, "
                        "%d agents, K = %d, M = 5, no map, goal noise 0.2, lockstep replans" % (nm, world, nm * world, N * world, Kn),
            "value": N * world * T / (ms_total * 1e-3), "unit": UNIT, "ms_per_step": ms_total / T, "steps": T, "warmup": W,
            "scaling": "weak", "missions": nm * world, "agents": N * world, "stages_ms": stage_ms, "qp_failsafe_agents": fails,
            "replay_exact": exact}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"           # keep NCCL's version banner off stdout: one JSON line only
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    W, K = max(args.warmup, 3), args.steps
    cfg, m = make_world(args)
    N = m.n_agents
    begin, NL = sharding.agent_block(N, world, rank)
    sl = slice(begin, begin + NL)

    rec, snap = pilot_rollout(cfg, m, args, dev.index, W + K)
    if snap["nbr_overflow"]:
        raise SystemExit("bench: %d agents had more than --max-nbr %d neighbours in range: their dropped neighbours would be "
                         "missing collision constraints (the reference has no cap); raise --max-nbr" % (snap["nbr_overflow"], args.max_nbr))
    edt = snap["edt"]
    pl = capi.SwarmPlanner(cfg, m, max_nbr=args.max_nbr, begin=begin, n_local=NL, device=dev.index)
    pl.build_edt(m.boxes)
    edt_ms = pl.edt_build_ms()
    pl.set_stream(torch.cuda.current_stream().cuda_stream)
    # per-step exchange of the agent records: peer-memory stores over NVLink (default) or an in-place NCCL all-gather
    exchange = sharding.RecordExchange(pl, world, rank, device=dev, mode=args.exchange)
    wp_dev = torch.from_numpy(np.ascontiguousarray(rec["wp"][:, sl])).to(dev)        # [T][NL][3] resident
    peak_fp64 = pl.measure_fp64_peak()

    gather = exchange.gather

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident: plan -> advance -> all-gather, everything on the device ----------------
    restore(pl, snap, sl)
    clocks = Clocks(dev.index)
    barrier()
    for t in range(W):
        pl.set_waypoints_device(wp_dev[t].data_ptr())
        pl.plan(); pl.advance(); gather()
    barrier()
    launches0 = pl.launch_count()
    clocks.begin()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for t in range(K):
        pl.set_waypoints_device(wp_dev[W + t].data_ptr())
        pl.plan(); pl.advance(); gather()
        ev[t + 1].record()
    barrier()
    clocks.end()
    launches = pl.launch_count() - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    total_ms = ev[0].elapsed_time(ev[K])
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    resident_final = pl.traj()
    replay_exact = bool(np.array_equal(resident_final, snap["final_traj"][sl]))

    # ---------------- per-kernel times: the same K steps again with CUDA events around every stage --------------
    # (the events serialise the SFC kernel, which otherwise runs beside neighbour search + LSC on a second stream,
    # so each kernel is timed alone, on the stream it is launched on)
    restore(pl, snap, sl)
    barrier()
    for t in range(W):
        pl.set_waypoints_device(wp_dev[t].data_ptr())
        pl.plan(); pl.advance(); gather()
    barrier()
    pl.enable_timing(True)
    clocks.begin()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for t in range(K):
        pl.set_waypoints_device(wp_dev[W + t].data_ptr())
        pl.plan(); pl.advance(); gather()
    es1.record()
    barrier()
    clocks.end()
    stage_ms, n_timed = pl.timings()
    pl.enable_timing(False)
    serial_ms = es0.elapsed_time(es1) / K

    # ---------------- work counters of the timed region (untimed replay, deterministic) ----------------
    restore(pl, snap, sl)
    flops, gjk_iters, edt_lookups, qp_iter_sum, pairs, sfc_alg, sfc_sat, sfc_tests = 0.0, 0, 0, 0, 0, 0, 0, 0
    nnz = pair_nnz(cfg)
    n_prof = min(K, 10)
    for t in range(W + n_prof):
        pl.set_waypoints_device(wp_dev[t].data_ptr())
        pl.plan()
        if t >= W:
            c = pl.counters()
            _, cnt = pl.neighbours()
            it = pl.qp_iters()
            flops += qp_flops(cfg, cnt, it, nnz)
            gjk_iters += c["gjk_iters"]; edt_lookups += c["edt_lookups"]; qp_iter_sum += c["qp_iters"]; pairs += c["pairs"]
            sfc_alg += c["sfc_vertices_alg"]; sfc_sat += c["sfc_tests_sat"]; sfc_tests += c["sfc_tests_mask"] + c["sfc_tests_records"]
        pl.advance(); gather()
    flops /= n_prof
    qp_ms = stage_ms["qp"]
    achieved_tf = flops / (qp_ms * 1e-3) / 1e12 if qp_ms > 0 else 0.0
    lsc_fl = lsc_flops(cfg, pairs / n_prof, gjk_iters / n_prof)
    sfc_bytes = 16.0 * sfc_alg / n_prof

    # ---------------- e2e: host buffers in, host buffers out, every step ----------------
    restore(pl, snap, sl)
    hb = {k: pinned(rec[k][:, sl].shape, torch.float32) for k in ("pos", "vel", "acc", "wp")}
    for k in hb:
        hb[k][1][...] = rec[k][:, sl]
    traj_t, traj_h = pinned((NL, cfg.M, cfg.n + 1, 3), torch.float32)
    structs = [capi.DlscAgents(hb["pos"][1][t].ctypes.data, hb["vel"][1][t].ctypes.data, hb["acc"][1][t].ctypes.data,
                               hb["wp"][1][t].ctypes.data, None) for t in range(W + K)]
    import ctypes as C

    # the trajectories go straight into the pinned host buffer (dlsc_bind_traj_host): every agent's result is written over
    # PCIe as soon as its QP finishes, so the device->host transfer of the step's result overlaps the QPs still running;
    # the buffer is complete after the step's dlsc_sync
    pl.bind_traj_host(traj_h)

    def e2e_step(t):
        pl.set_agents_async(structs[t])
        gather()
        pl.plan()
        pl.publish_records()
        pl.sync()                                                                 # result of this step is in traj_h

    barrier()
    for t in range(W):
        e2e_step(t)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.begin()
    t_host0 = time.perf_counter()
    e0.record()
    for t in range(K):
        e2e_step(W + t)
    e1.record()
    t_host = (time.perf_counter() - t_host0) * 1e3        # the last dlsc_get_traj of the loop has synchronised the stream
    barrier()
    clocks.end()
    clk = clocks.stop()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    tt = torch.tensor([e2e_ms, t_host], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms, t_host = float(tt[0].item()), float(tt[1].item())
    e2e_exact = bool(np.array_equal(traj_h, snap["final_traj"][sl]))
    pl.bind_traj_host(None)
    h2d = int(4 * N * 12)
    d2h = int(N * cfg.M * (cfg.n + 1) * 12)

    # ---------------- closed loop with the waypoint provider (1 GPU): every step, host side ----------------
    # D2H states + trajectories -> dlsc_wp_step (comm-range groups + PIBT + update rules, the reference's own per-step host
    # stage, src/multi_sync_simulator.cpp:308-466) -> H2D waypoints -> dlsc_step -> dlsc_advance.  Reported beside `e2e`
    # (which replays recorded waypoints): at 4096 agents the reference's waypoint algorithm itself costs ~0.4 s per step on a
    # host core (it replans every agent's whole lattice path to completion each step), far more than the batched replan.
    closed = None
    if world == 1 and args.closed_loop_steps > 0:
        restore(pl, snap, sl)
        wp = capi.WaypointProvider(cfg, m, edt=(edt[0], edt[1], edt[2], edt[3], cfg.world_res))
        wcur = rec["wp"][0].copy()
        traj_cl = None
        t_wp, t_all = [], []
        for t in range(args.closed_loop_steps + 1):
            t0 = time.perf_counter()
            pos, _, _ = pl.state()
            goal_cur = pl.goal() if t else snap["records"][:, cfg.M * (cfg.n + 1) * 3 + 6: cfg.M * (cfg.n + 1) * 3 + 9].copy()
            t1 = time.perf_counter()
            wcur = wp.step(pos, goal_cur, traj_cl, wcur)
            t2 = time.perf_counter()
            pl.set_agents(waypoint=wcur)
            pl.plan()
            traj_cl = pl.traj()
            pl.advance(); pl.sync()
            t3 = time.perf_counter()
            if t:                                      # the first call also builds the provider's distance tables
                t_wp.append(t2 - t1); t_all.append(t3 - t0)
        closed = {"steps": len(t_all), "ms_per_step": 1e3 * float(np.mean(t_all)), "waypoint_provider_ms_per_step": 1e3 * float(np.mean(t_wp)),
                  "pibt_timesteps_last": wp.pibt_timesteps(), "value": N / float(np.mean(t_all)), "unit": UNIT,
                  "what": "host clock; per step: states + trajectories D2H, dlsc_wp_step on one host core (comm-range groups, PIBT to "
                          "completion for all %d agents, update rules), waypoints H2D, dlsc_step, dlsc_advance" % N}
        wp.close()

    # ---------------- dynamic obstacles (SURVEY s8(f4), 1 GPU): the same swarm with obstacles flying through it ----------------
    # every step: dlsc_set_obstacles (host states in) -> dlsc_step (obstacle prediction, k_lsc_dyn, k_trap, slack QP: active set,
    # interior point for the slack-heavy agents) -> dlsc_advance.  Obstacles: 0.3 m spheres crossing the swarm area at 1 m/s, max_acc 2,
    # opt/slack_collision_weight 100 (launch/testall_DLSCGC_3D.launch).  Reported beside the swarm-only headline.
    dyn = None
    if world == 1 and args.dyn_obstacles > 0:
        restore(pl, snap, sl)
        nd = min(args.dyn_obstacles, 16)
        rng = np.random.default_rng(1234)
        he = float(m.world_max[0]) * 0.7
        opos = np.stack([rng.uniform(-he, he, nd), rng.uniform(-he, he, nd), np.full(nd, 1.0)], axis=1).astype(np.float32)
        ang = rng.uniform(0, 2 * np.pi, nd)
        ovel = np.stack([np.cos(ang), np.sin(ang), np.zeros(nd)], axis=1).astype(np.float32)
        kw = dict(radius=0.3, downwash=1.0, max_acc=2.0, slack_weight=100.0)
        occupied = missions.occupied_nodes(m.boxes, cfg.grid_res)
        goal_des = m.goal.astype(np.float32)
        wpc = rec["wp"][0].copy()                       # waypoints of the restored state
        traj_d, dev_ms, host_ms, fails_total = None, [], [], 0
        n_dw = 3
        for t in range(n_dw + args.dyn_steps):
            # waypoints closed loop on the host, untimed (agents displaced by an obstacle must not be dragged along by
            # waypoints recorded without it: the communication-range rows tie every trajectory to its waypoint)
            pos, _, _ = pl.state()
            if t > 0:
                wpc = missions.next_waypoints(wpc, pl.goal(), goal_des, traj_d, pos, cfg, occupied)
            pl.set_agents(waypoint=wpc)
            pl.sync(); torch.cuda.synchronize()
            ed0, ed1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            th0 = time.perf_counter()
            ed0.record()
            pl.set_obstacles(opos, ovel, **kw)
            pl.plan(); pl.advance()
            ed1.record(); pl.sync(); torch.cuda.synchronize()
            th1 = time.perf_counter()
            traj_d = pl.traj()
            opos = opos + ovel * np.float32(cfg.dt)
            if t >= n_dw:
                dev_ms.append(ed0.elapsed_time(ed1)); host_ms.append(1e3 * (th1 - th0))
                fails_total += int(((pl.status() & capi.FAIL_MASK & ~capi.NBR_OVERFLOW) != 0).sum())
        st = pl.status(); sk = pl.slack()
        cpu_dyn = None
        if not args.no_cpu_baseline:
            # the CPU oracle (test infrastructure, here only as the timed baseline of this leg) on the next replan of the same state
            snap_d = {"records": pl.get_records(), "sfc": pl.sfc(), "acc": pl.state()[2], "seq": pl.seq}
            sw = oracle_swarm(cfg, m, edt, snap_d, {"wp": [wpc]}, args, os.cpu_count() or 1)
            sw.set_obstacles(opos, ovel, **kw)
            tc0 = time.perf_counter()
            sw.step()
            cpu_dyn = {"ms_per_step": 1e3 * (time.perf_counter() - tc0), "cores": os.cpu_count() or 1, "kind": "port",
                       "sample": "one replan of all %d agents with the same %d obstacles, oracle on all host threads" % (N, nd)}
        dyn = {"obstacles": nd, "steps": args.dyn_steps, "ms_per_step": float(np.mean(host_ms)), "p50_step_ms": float(np.median(host_ms)),
               "device_ms_per_step": float(np.mean(dev_ms)), "value": N / (1e-3 * float(np.mean(host_ms))), "unit": UNIT,
               "qp_failsafe_agent_steps": fails_total, "cpu_baseline": cpu_dyn,
               "last_step": {"nbr_overflow_agents": int(((st & capi.NBR_OVERFLOW) != 0).sum()),
                             "agents_on_interior_point": int(((st & capi.QP_IPM_USED) != 0).sum()),
                             "agents_using_slack": int((sk.min(axis=(1, 2)) < -1e-6).sum()), "min_slack_m": float(sk.min())},
               "what": "host clock around dlsc_set_obstacles + dlsc_step + dlsc_advance, per step, states device-resident; waypoints "
                       "generated closed loop on the host between the timed regions; %d obstacles (r 0.3 m, 1 m/s straight lines, "
                       "max_acc 2), slack weight 100.  The step is set by the few agents beside an obstacle: their QPs keep tens of "
                       "rows active and go to the interior-point kernel" % nd}

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tot_stage = sum(stage_ms.values())
        out = {
            "metric": METRIC, "value": N * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "p50_step_ms": statistics.median(step_ms), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg, m),
            "exchange": {"p2p": "peer-memory all-gather over NVLink (dlsc_exchange_records), %d-float records" % pl.rec_floats,
                         "nccl": "NCCL all-gather", "single": "no exchange (1 GPU)"}.get(exchange.mode, exchange.mode),
            # e2e: host wall clock around the K steps (host buffers in, host buffers out, every step); the CUDA-event time of
            # the same region is reported beside it
            "e2e": {"value": N * K / (t_host * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_host / K, "device_ms_per_step": e2e_ms / K, "clock": "host perf_counter, max over ranks",
                    "replay_exact": e2e_exact},
            "gpu_launches": int(launches),
            "roofline": None,
            "stages_ms": stage_ms, "stages_note": "second pass over the same %d steps with CUDA events around every stage "
            "(stages serialised: %.4f ms/step; the headline pass overlaps k_sfc with k_neighbours + k_lsc on two streams)" % (
                K, serial_ms), "qp_share": qp_ms / tot_stage if tot_stage > 0 else None,
            "work_per_step": {"pairs": pairs / n_prof, "gjk_iters": gjk_iters / n_prof, "sfc_vertices": edt_lookups / n_prof,
                              "sfc_vertices_algorithmic": sfc_alg / n_prof, "sfc_box_tests": sfc_tests / n_prof,
                              "sfc_box_tests_sat": sfc_sat / n_prof, "qp_iters": qp_iter_sum / n_prof},
            "pilot": {"qp_failsafe_agents": snap["fails"], "nbr_overflow_agents": snap["nbr_overflow"],
                      "mean_dist_to_goal_m": snap["dist_to_goal"], "replay_exact": replay_exact},
            "clocks": clk,
        }
        if closed:
            out["closed_loop"] = closed
        if dyn:
            out["dynamic_obstacles"] = dyn
        hbm = peaks.get("hbm_gbs") or 6650.0
        ncell = int(edt[2][0]) * int(edt[2][1]) * int(edt[2][2])
        out["edt_build"] = {"kernels": "k_edt_pass_z_cols + k_edt_col_any/window + k_edt_pass_y/x", "ms": edt_ms, "cells": ncell, "bound": "hbm",
                            "achieved": 17.0 * ncell / (edt_ms * 1e-3) / 1e9 if edt_ms > 0 else None, "peak": hbm, "unit": "GB/s",
                            "frac": 17.0 * ncell / (edt_ms * 1e-3) / 1e9 / hbm if edt_ms > 0 else None,
                            "algorithmic": "once per mission: 1 B occupancy in + 16 B record out per cell"}
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass

        def rl(kernel, bound, work, ms, peak, unit, scale, extra):
            a = work / (ms * 1e-3) / scale if ms > 0 else 0.0
            d = {"kernel": kernel, "bound": bound, "achieved": a, "peak": peak, "unit": unit, "frac": a / peak if peak else None,
                 "traffic": traffic.get(kernel), "work_per_launch": work, "kernel_ms": ms}
            d.update(extra)
            return d
        # LSC: SURVEY s8(d) gives bytes AND flops per agent; with the measured GJK depth (~1.5 iterations per hull) the
        # intensity is ~1.5 flop/B, far below the machine balance (36 TFLOP/s : 6.4 TB/s), so the HBM roof is the one that
        # bounds it; the FP64 fraction is reported beside it.  Bytes: read (K+1) M P 12 + K 36, write K M (12 + 8 P + 4).
        Mq, Pq = cfg.M, cfg.n + 1
        lsc_bytes = (pairs / n_prof + NL) * Mq * Pq * 12.0 + (pairs / n_prof) * 36.0 + (pairs / n_prof) * Mq * (12.0 + 8.0 * Pq + 4.0)
        sfc_traffic = traffic.get("k_sfc")
        rls = {
            "sfc": rl("k_sfc", "latency", sfc_traffic or 0.0, stage_ms["sfc"], hbm, "GB/s", 1e9, {
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks.get("hbm_gbs") else "fallback",
                "algorithmic": "achieved = the kernel's measured DRAM traffic (profiles/traffic.json, ncu dram__bytes) / its "
                               "CUDA-event time: the kernel answers %.1f%% of its box tests with an 8-corner summed-area query and "
                               "the rest from a 1-byte-per-vertex mask, so SURVEY s8(d)'s 16 B x lattice-vertex figure "
                               "(%.0f MB per step here) is not what it moves" % (100.0 * sfc_sat / max(sfc_tests, 1), sfc_bytes / 1e6),
                "note": "bound by the serial greedy chain (~%d dependent box tests per agent, each a few hundred scalar "
                        "instructions of control arithmetic), not by HBM or a pipe" % round(sfc_tests / n_prof / max(NL, 1))}),
            "lsc": rl("k_lsc+k_lsc_rest", "hbm", lsc_bytes, stage_ms["lsc"], hbm, "GB/s", 1e9, {
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks.get("hbm_gbs") else "fallback",
                "fp64": {"achieved": lsc_fl / (stage_ms["lsc"] * 1e-3) / 1e12 if stage_ms["lsc"] > 0 else 0.0, "peak": peak_fp64,
                         "unit": "TFLOP/s", "frac": (lsc_fl / (stage_ms["lsc"] * 1e-3) / 1e12 / peak_fp64) if stage_ms["lsc"] > 0 and peak_fp64 else None,
                         "work_per_launch": lsc_fl}}),
            "qp": rl("k_qp_fast+k_qp_gi", "fp64", flops, qp_ms, peak_fp64, "TFLOP/s", 1e12, {
                "peak_source": "dlsc_measure_fp64_peak (DFMA chains, measured in this run; MEASURED_PEAKS.json has no FP64 figure)"}),
        }
        rls["lsc"]["traffic"] = (traffic.get("k_lsc") or 0) + (traffic.get("k_lsc_rest") or 0) or None
        dom = max(("lsc", "qp"), key=lambda k: stage_ms[k])      # the dominant kernel with a throughput roof (k_sfc is latency bound)
        out["roofline"] = rls[dom]
        out["rooflines_other"] = {k: v for k, v in rls.items() if k != dom}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(cfg, m, edt, rec, snap, args, steps=1)
    if world > 1:
        if exchange.mode == "p2p":
            pl.p2p_status()                      # raises if a peer ever missed an exchange
            pl.p2p_disconnect()
        dist.barrier()
    pl.close()
    if args.mc_missions > 0:
        mc = run_montecarlo(args, dev, world, rank, dist)
        if rank == 0:
            out["montecarlo"] = mc
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------
def oracle_params_of(cfg, m):
    from oracle import oracle_py as O
    return O.make_params(M=cfg.M, n=cfg.n, phi=cfg.phi, dim=cfg.dim, use_sfc=cfg.use_sfc, dt=cfg.dt,
                         world_min=m.world_min, world_max=m.world_max, world_res=cfg.world_res, grid_res=cfg.grid_res,
                         z_2d=cfg.z_2d, comm_range=cfg.comm_range, w_control=cfg.w_control, w_terminal=cfg.w_terminal,
                         reset_threshold=cfg.reset_threshold)


def oracle_swarm(cfg, m, edt, snap, rec, args, n_threads):
    """The CPU oracle (oracle/: test infrastructure, here only as the timed CPU baseline) loaded with the
    planner state at the start of the timed region."""
    from oracle import oracle_py as O
    O.build()
    p = oracle_params_of(cfg, m)
    e = O.Edt(p, edt[0], edt[1], edt[2], edt[3])
    sw = O.Swarm(p, m.start, m.goal, m.radius, m.downwash, m.max_vel, m.max_acc, m.nominal_vel, edt=e,
                 max_nbr=args.max_nbr, n_threads=n_threads)
    N, o = m.n_agents, cfg.M * (cfg.n + 1) * 3
    r = snap["records"]
    sw.traj[...] = r[:, :o].reshape(sw.traj.shape)
    sw.pos[...] = r[:, o:o + 3]; sw.vel[...] = r[:, o + 3:o + 6]; sw.goal_cur[...] = r[:, o + 6:o + 9]
    sw.acc[...] = snap["acc"]; sw.sfc[...] = snap["sfc"]; sw.sfc_init[...] = 0
    sw.waypoint[...] = rec["wp"][0]
    sw.seq = snap["seq"]
    return sw


def _timed_replans(cfg, m, edt, snap, rec, args, n_threads, seconds, n_cap=None):
    """Repeat the replan of agents [0, n) from the snapshot state until ~`seconds` of wall time are timed."""
    N = m.n_agents
    sw = oracle_swarm(cfg, m, edt, snap, rec, args, n_threads)
    probe = min(N, 4 * n_threads)
    seq0 = sw.seq
    t0 = time.perf_counter(); sw.step(0, probe); t_probe = time.perf_counter() - t0
    sw.seq = seq0
    n = int(min(N, max(probe, seconds / max(t_probe / probe, 1e-9))))
    n = max(n_threads, (n // n_threads) * n_threads)
    if n_cap:
        n = min(n, n_cap)
    t, reps, stage = 0.0, 0, None
    while reps == 0 or (t < seconds and reps < 64):
        sw = oracle_swarm(cfg, m, edt, snap, rec, args, n_threads)
        t0 = time.perf_counter(); sw.step(0, n); t += time.perf_counter() - t0
        reps += 1
        ss = np.array([float(x) for x in sw.stage_seconds])
        stage = ss if stage is None else stage + ss
    ok = int(((sw.status[:n] & FAIL_BITS) == 0).sum())
    return n, reps, t, stage, ok, sw


FAIL_BITS = 1 | 2 | 4 | 8     # QP_MAXITER | QP_NUMERIC | SFC_INIT_FAILED | GOAL_INFEASIBLE (include/dlsc_b200.h)


def highs_qp_leg(cfg, m, sw, n_qp=16):
    """BASELINE.md s3 leg 3: the QP stage of sample agents solved by an independent CPU solver (HiGHS through scipy, the
    x-space restatement of tests/qp_highs.py) -- mean seconds per solve, build time excluded."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import qp_highs
        times, status = [], []
        for a in range(min(n_qp, m.n_agents)):
            K = int(sw.nbr_cnt[a])
            qp = qp_highs.build_qp(cfg.M, cfg.n, cfg.dim, cfg.dt, cfg.w_control, cfg.w_terminal, m.world_min, m.world_max,
                                   cfg.comm_range, sw.pos[a], sw.vel[a], sw.acc[a], sw.goal_cur[a], sw.waypoint[a], sw.radius[a],
                                   sw.max_vel[a], sw.max_acc[a], sw.nominal_vel[a], sfc=sw.sfc[a] if cfg.use_sfc else None,
                                   lsc_normal=sw.lsc_normal[a, :K], lsc_anchor=sw.lsc_anchor[a, :K], lsc_d=sw.lsc_d[a, :K])
            t0 = time.perf_counter()
            _, _, st = qp_highs.solve_highs(qp, time_limit=5)
            times.append(time.perf_counter() - t0); status.append(st)
        return {"ms_per_qp": 1e3 * float(np.mean(times)), "qps": len(times), "optimal": int(sum(x == "Optimal" for x in status)),
                "what": "HiGHS convex-QP solve (1 thread) of the first %d agents' QPs at the first timed step, rows built from the "
                        "same LSC / SFC constraints; model build excluded" % len(times)}
    except Exception as e:      # scipy without the private HiGHS binding
        return {"unavailable": repr(e)[:200]}


def cpu_baseline(cfg, m, edt, rec, snap, args, steps=1):
    """The three CPU legs BASELINE.md s3 promises: all host threads (one agent per thread), one thread (the reference's
    real execution model, src/multi_sync_simulator.cpp:516-524), and HiGHS on the QP stage."""
    cores = os.cpu_count() or 1
    N = m.n_agents
    n, reps, t, stage, ok, sw = _timed_replans(cfg, m, edt, snap, rec, args, cores, args.cpu_seconds)
    n1, reps1, t1, stage1, ok1, _ = _timed_replans(cfg, m, edt, snap, rec, args, 1, min(6.0, args.cpu_seconds))
    return {"value": n * reps / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "agents [0,%d) of the same %d-agent swarm at the first timed step, one replan each, one agent per "
                      "thread on %d threads, repeated %d times from the same state (oracle/ C++ port; the reference itself "
                      "needs ROS+CPLEX and cannot run here); %d/%d QPs converged; %.1f s timed" % (n, N, cores, reps, ok, n, t),
            "stage_seconds": [float(x) for x in stage],
            "serial_1_thread": {"value": n1 * reps1 / t1, "unit": UNIT, "cores": 1, "ms_per_agent_replan": 1e3 * t1 / (n1 * reps1),
                                "sample": "agents [0,%d), %d passes, %.1f s timed; stage seconds (predict+nbr, LSC, SFC, goal, QP): %s" % (
                                    n1, reps1, t1, ", ".join("%.3f" % x for x in stage1))},
            "highs_qp": highs_qp_leg(cfg, m, sw),
            "reference_logged": {"ms_per_agent_replan": 12.17, "unit": "ms", "what": "the reference's own log (maze10_dense, 10 agents, "
                                 "2-D, CPLEX, unknown CPU): log/summary_DLSCGC_10agents.csv planning_time_average"}}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on the box's host cores.  The literal ROS + CPLEX build
    cannot exist in this image, so it is the oracle port (oracle/), one agent per thread on every host thread.  It never
    loads libdlsc_b200.so: the swarm is settled with the oracle itself (same seeded world, same waypoint provider), then
    W + K successive replan steps of the whole swarm (or of a bounded agent sample when a step would take too long) are
    timed with the host clock."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O
    W, K = max(args.warmup, 1), args.steps
    cfg, m = make_world(args)
    N = m.n_agents
    cores = os.cpu_count() or 1
    O.build()
    p = oracle_params_of(cfg, m)
    e = O.edt_build(p, m.boxes)
    sw = O.Swarm(p, m.start, m.goal, m.radius, m.downwash, m.max_vel, m.max_acc, m.nominal_vel, edt=e,
                 max_nbr=args.max_nbr, n_threads=cores)
    occupied = missions.occupied_nodes(m.boxes, cfg.grid_res)
    goal_des = m.goal.astype(np.float32)
    t_start = time.perf_counter()
    n = N                                                # every step replans the whole swarm: a consistent rollout
    traj = None
    times, fails = [], 0
    for s in range(args.settle + W + K):
        sw.waypoint = missions.next_waypoints(sw.waypoint, sw.goal_cur, goal_des, traj, sw.pos, cfg, occupied)
        t0 = time.perf_counter()
        st = sw.step()
        dt = time.perf_counter() - t0
        traj = sw.traj.copy()
        sw.advance()
        if s >= args.settle + W:
            times.append(dt)
            fails += int(((st & FAIL_BITS) != 0).sum())
    total = sum(times)
    val = n * K / total
    sample = ("%d settle + %d warm-up + %d timed successive steps of the %d-agent swarm, %s replanned per step, one agent per "
              "thread on %d host threads (oracle/ C++ port: LSC with the openGJK restatement, SFC, goal, dense interior-point "
              "QP); %d QP failsafes; %.1f s wall in total" % (
                  args.settle, W, K, N, "all agents", cores, fails, time.perf_counter() - t_start))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": total / K * 1e3 * (N / n), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, cfg, m),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "kind \"port\": the reference's own build needs ROS, octomap, dynamicEDT3D and CPLEX (absent, no network); "
                "ms_per_step is extrapolated to the full swarm when a sample was replanned",
    }))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
