#!/usr/bin/env python
"""bench.py -- MPC QP solves/s of the B200 engine on BASELINE.json's configs[1]:
"Batched Solo trot N=16, 4096 independent instances on 1 B200 with warm start across ticks".

A step = one MPC tick of the whole batch: build + solve + extract for B robots (per GPU).
    value   device-resident: the (W+K) ticks of closed-loop inputs already sit in HBM; CUDA events
            on the engine's stream around the K timed ticks, max over ranks.
    e2e     the same K ticks through the public host API (pinned host xref/fsteps in, forces out),
            H2D + D2H inside the timed region.
Inputs are produced untimed by running the engine itself in closed loop on the scenario generator
(mpc-tsid_b200/scenario.py), then replayed: the solver is deterministic, so the replay reproduces the
closed loop, warm starts included.  Multi-GPU: instances shard by index, no collective on the hot
path ("weak": every GPU gets its own B robots); torch.distributed is used only for the barrier and
the max-over-ranks of the timings.

`--impl reference` times the reference's CPU path instead: since `osqp` is not installable here and
/root/reference does not travel to the GPU box, that is the oracle port in plain C (oracle/mpc_osqp.c at eps 1e-8),
one robot per host thread.
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
for p in (ROOT, os.path.join(ROOT, "mpc-tsid_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "MPC QP solves/sec (Solo trot, N=16)"
WORKLOAD = ("BASELINE.json configs[1]: batched Solo trot N=16 dt=0.02, %d instances per GPU, closed loop with warm start "
            "across ticks, random commands/gait phases, seeded state noise")
UNIT = "solves/s"
N_STEPS = 16


# ----------------------------------------------------------------------------------------------
# CPU arm: the plain-C restatement of MPC.py + the OSQP algorithm (oracle/mpc_osqp.c), one robot per
# host thread.  Closed-loop input sequences are recorded first (untimed, the oracle itself in the
# loop), then replayed by `cores` threads at once: the timed region holds nothing but the path
# (build -> warm start -> update -> solve -> extract), tick after tick, exactly as MPC.run does.
# ----------------------------------------------------------------------------------------------
def _record_closed_loop(args):
    wid, ticks = args
    from oracle import c_port
    from scenario import Scenario
    sc = Scenario(1, gaits="trot", seed=20260 + 1000 + wid)
    m = c_port.MPC(n_steps=N_STEPS, eps=1e-8)
    xs, fs = [], []
    for t in range(ticks):
        xref, fsteps = sc.inputs()
        r = m.run(t == 0, xref[0], fsteps[0])                                   # MPC.py:460-514
        xs.append(xref[0]); fs.append(fsteps[0])
        x_robot = r["x"][:12 * N_STEPS].reshape((12, N_STEPS), order="F") + xref[0][:, 1:]    # MPC.py:437
        sc.advance(x_robot[:, 0][None])
    m.close()
    return np.stack(xs), np.stack(fs)


def cpu_arm(steps, warmup, cores=None):
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_port
    c_port.load()
    cores = cores or os.cpu_count() or 1
    distinct = min(cores, 16)
    T = warmup + steps
    with ThreadPoolExecutor(distinct) as pool:          # ctypes releases the GIL inside the C solve
        recs = list(pool.map(_record_closed_loop, [(w, T) for w in range(distinct)]))
    xr = np.stack([recs[i % distinct][0] for i in range(cores)])
    fs = np.stack([recs[i % distinct][1] for i in range(cores)])
    wall, f_ref, iters = c_port.replay_mt(xr, fs, warm=warmup, n_steps=N_STEPS, eps=1e-8)
    # the same replay with the cost scale OSQP would pick if it left the (all-zero) q out of its rule: context for the ratio, since
    # the faithful rule costs ~20x more ADMM iterations on this QP (tools/baseline_experiments.py)
    c_port.set_cost_scaling_variant(True)
    try:
        wall_v, f_var, iters_v = c_port.replay_mt(xr, fs, warm=warmup, n_steps=N_STEPS, eps=1e-8)
    finally:
        c_port.set_cost_scaling_variant(False)
    return dict(value=cores * steps / wall, unit=UNIT, cores=cores, kind="port",
                osqp_iterations_per_solve=iters, eps=1e-8,
                cost_scaled_variant={"value": cores * steps / wall_v, "unit": UNIT, "osqp_iterations_per_solve": iters_v,
                                     "max_force_difference_N": float(np.abs(f_ref[:, warmup:] - f_var[:, warmup:]).max()),
                                     "what": "same replay, cost scale c = 1 / mean column norm of P instead of OSQP's published rule "
                                             "(which treats the reference's q = 0 as 1 and leaves c = 1): NOT the reference's behaviour, "
                                             "shown because the GPU / CPU ratio moves ~20x with it"},
                sample="%d host threads x %d consecutive closed-loop trot ticks each (after %d untimed warm-up ticks; %d distinct "
                       "robots) through oracle/mpc_osqp.c = C restatement of MPC.py's build + the OSQP algorithm (sparse LDL', "
                       "Ruiz scaling, adaptive rho, warm start) at eps 1e-8, polish off; the reference's own build half is "
                       "Python (~1.3 ms per tick, SURVEY.md 6) and would be slower than this" % (cores, steps, warmup, distinct)), wall


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base, wall = cpu_arm(args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % 4096,
                   "sample": "one closed-loop trot robot per host core, one QP per tick (same generator, same tick shape)"},
        "cpu_baseline": base, "cores": base["cores"],
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "osqp is not installable offline and /root/reference is absent on the GPU box: this is the oracle port in C "
                "(oracle/mpc_osqp.c: restated MPC.py build + restated OSQP algorithm), a step = one tick of every host thread",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# FLOP model of what the engine actually executes, per instance-tick
# ----------------------------------------------------------------------------------------------
def flops_per_tick(sweeps, iters, fallbacks, n=96, N=N_STEPS, refine=0):
    """DENSE path (mode 3), per instance-tick, counted from the algorithm (not from SASS):
    sweep      = Cholesky n^3/3 + triangular inverse n^3/3 (DMMA) + block assembly
                 + (1 + refine) solves (two triangular mat-vecs, 2 n^2) + (1 + refine) Hessian applies
    admm setup = the same factorisation once;  admm iteration = one solve + per-foot projection."""
    feet = 4 * N
    fact = 2.0 * n ** 3 / 3.0
    assemble = feet * 21 * 7
    solve = 2.0 * n * n + 2 * feet * 36
    grad = 2.0 * 2 * 6 * N * N + 2 * feet * 36
    sweep = fact + assemble + (1 + refine) * solve + (1 + refine) * grad
    admm_setup = fact + assemble
    admm_iter = solve + feet * 60
    fixed = feet * 60 + 6 * N * N * 8 + 12 * N * 8
    return sweeps * sweep + fallbacks * admm_setup + iters * admm_iter + fixed


def flops_stagewise_sweep(N=N_STEPS):
    """STAGE-WISE path (default), one active-set sweep of one robot, counted from the algorithm of
    mpcqp_riccati.cuh (useful work only: the redundant per-lane copies of the 6x6 factorisations are not counted).
    Per stage: two 6x6 factorisations (2 x n^3/3; the capacity-64 kernel, N > 32, also forms the inverses of the two unit
    triangular factors, another 2 x n^3/3 -- capacities 16 and 32 substitute with the factors themselves), T = E L and
    G = I + L'T (triangular / symmetric products), 13 rows x 4 triangular row products or substitutions (21 FMA each),
    12 rows x (36 + 6) FMA for Pt[:, p] and pt, 12 x 6 FMA for Pt[:, v] beta, the assembly of P_k."""
    n = 6
    chol = 2 * (n ** 3 / 3.0) * (2 if N > 32 else 1)
    t_g = 2 * 126 + 2 * 56
    rows = 13 * 4 * 21 * 2
    wpart = 12 * (36 + 6) * 2 + 12 * 6 * 2
    assemble_p = 21 * 2 + 36 * 2 + 21 * 5 + 12 * 3
    stage = chol + t_g + rows + wpart + assemble_p
    feet = 4 * N
    e_beta = feet * (18 + 8 + 45 + 66 + 15)          # lever block, S = Z R^-1 Z', A S, the 21 entries, impulse of pf
    forward = N * (12 * 2 + 6 * 13 * 2 + 6)
    costate = N * 6 * 6
    per_foot = feet * (18 + 18 + 12 + 12 + 20)       # lever block, Bv' lam, face solve, gradient, guard
    return N * stage + e_beta + forward + costate + per_foot


def flops_stagewise_fixed(N=N_STEPS):
    feet = 4 * N
    return feet * 40 + N * 30 + 12 * N * 4 + feet * 40   # decode, inertia blocks, outputs/objective, multipliers


def profile_metric(path, name):
    """A `name [unit]: value` line of a committed ncu summary (profiles/*.txt written by tools/summarize_ncu.py)."""
    try:
        for ln in open(os.path.join(ROOT, path)):
            if ln.startswith(name + " ["):
                unit = ln[ln.index("[") + 1:ln.index("]")]
                v = float(ln.split(":")[-1].split("|")[0].strip().replace(",", ""))
                return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    except (OSError, ValueError):
        pass
    return None


def run_sweep(torch, mpcqp, Scenario, rank, local_rank, world, barrier, max_over_ranks, B, K, W, gaits, n_steps=N_STEPS, mode=None,
              e2e=True, peak_tflops=None):
    """Device-resident closed loop (BASELINE configs[2..4] shapes): every tick plans footsteps and the reference trajectory,
    builds and solves the QP and integrates the centroidal state, per robot, inside the solve kernel (mpcqp_scenario_run).
    `value` times K ticks back to back; `e2e` runs the same ticks one C-ABI call at a time and reads every robot's state back."""
    sc = Scenario(B, n_steps=n_steps, gaits=gaits, seed=4242 + rank, noise_kind="hash")
    kw = {} if mode is None else {"mode": mode}
    eng = mpcqp.Engine(batch=B, n_steps=n_steps, device=local_rank, **kw)
    eng.scenario_init(sc)
    eng.scenario_run(W)
    eng.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = eng.launches
    e0.record(stream)
    eng.scenario_run(K)
    e1.record(stream)
    eng.synchronize()
    barrier()
    launches = eng.launches - l0
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    info = eng.info(with_y=False)
    out = {"value": world * B * K / (total_ms * 1e-3), "unit": UNIT, "ms_per_tick": total_ms / K, "ticks": K, "instances_per_gpu": B,
           "n_steps": n_steps, "gaits": "/".join(gaits), "gpu_launches": int(launches),
           "unsolved_instances_last_tick": int((info["status"] != 1).sum()),
           "sweeps_per_solve_last_tick": float(info["sweeps"].mean()), "fallback_frac_last_tick": float((info["iters"] > 0).mean())}
    if peak_tflops:
        flop = float(info["sweeps"].mean()) * flops_stagewise_sweep(n_steps) + flops_stagewise_fixed(n_steps)
        out["roofline_frac"] = flop * B * K / (total_ms * 1e-3) * 1e-12 / peak_tflops
    if e2e:
        barrier()
        w0 = time.perf_counter()
        for _ in range(K):
            eng.scenario_run(1)
            st = eng.scenario_state()
        e2e_s = max_over_ranks(time.perf_counter() - w0)
        out["e2e"] = {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": B * 23 * 8,
                      "ms_per_step": 1e3 * e2e_s / K,
                      "what": "one mpcqp_scenario_run(1) + mpcqp_scenario_get per tick (states, frames, feet to the host)"}
        out["state_checksum"] = float(np.abs(st["state"]).sum())
    eng.close()
    return out


def sweep_main(args, torch, mpcqp, Scenario, rank, local_rank, world, dist, barrier, max_over_ranks, sum_over_ranks):
    """`--workload sweep | mixed-sweep`: the closed-loop scenario sweep as the bench line (BASELINE configs[4] / configs[2] shape)."""
    B, K, W = args.batch, args.steps, max(args.warmup, 3) + max(args.settle, 0)
    gaits = ["trot"] if args.workload == "sweep" else ["trot", "pace", "bound", "walk"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    r = run_sweep(torch, mpcqp, Scenario, rank, local_rank, world, barrier, max_over_ranks, B, K, W, gaits, n_steps=args.n_steps, mode=args.mode)
    clocks = sampler.stop()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC.replace("Solo trot", "closed-loop sweep").replace("N=16", "N=%d" % args.n_steps), "value": r["value"], "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_tick"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[4] shape: device-resident closed-loop sweep, %d robots per GPU, gaits %s, N=%d: footstep "
                                   "planner + QP build + solve + centroidal state integration per tick inside the solve kernel, seeded "
                                   "counter-based state noise" % (B, "/".join(gaits), args.n_steps),
                       "instances_per_gpu": B, "settle_ticks": max(args.settle, 0), "l2": "no inputs: the planner runs in the kernel; carried state %.0f MB" % (B * 6.5e-3)},
            "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": clocks,
            "unsolved_instances_last_tick": r["unsolved_instances_last_tick"],
            "sweeps_per_solve_last_tick": r["sweeps_per_solve_last_tick"], "fallback_frac_last_tick": r["fallback_frac_last_tick"],
            "state_checksum": r["state_checksum"]}))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="robots per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="leave the process on whatever cores the launcher gave it")
    ap.add_argument("--mode", type=int, default=13, help="solver stages (include/mpcqp.h MPCQP_MODE_*): 13 = stage-wise active set + interior-point "
                    "fallback (default), 7 = stage-wise + dense ADMM fallback, 3 = dense")
    ap.add_argument("--n-steps", type=int, default=N_STEPS, help="horizon of the sweep workloads (the trot headline is N = 16)")
    ap.add_argument("--latency-ticks", type=int, default=200, help="ticks of the separate latency window (SURVEY 8d: p50 / p99 over >= 200 ticks)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short sub-runs of BASELINE configs[2], [3], [4]")
    ap.add_argument("--no-dropin", action="store_true", help="skip the MPC_Wrapper end-to-end leg")
    ap.add_argument("--overlap", type=int, default=0, help="index ranges the device-resident ticks are issued as (mpcqp_set_overlap): "
                    "1 = one tick at a time, 0 = 8 ranges up to 6144 robots per GPU (4 when four or more GPU processes share the host), 2 beyond")
    ap.add_argument("--cpu-ticks", type=int, default=100)
    ap.add_argument("--workload", default="trot", choices=["trot", "sweep", "mixed-sweep"],
                    help="trot = BASELINE configs[1] (the headline line, default); sweep / mixed-sweep = device-resident closed-loop "
                         "scenario sweep (configs[4] shape: planner + QP + state integration per tick on the GPU, trot or mixed gaits)")
    ap.add_argument("--settle", type=int, default=20,
                    help="closed-loop ticks run (untimed) before the warm-up so that the timed ticks are steady-state "
                         "operation, not the cold-start transient of robots released from rest")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)

    import torch
    import mpcqp
    from scenario import Scenario

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    affinity = mpcqp.bind_near_gpu(local_rank) if not args.no_bind else {"off": True}    # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if args.workload != "trot":
        return sweep_main(args, torch, mpcqp, Scenario, rank, local_rank, world, dist, barrier, max_over_ranks, sum_over_ranks)

    B, N, K = args.batch, N_STEPS, args.steps
    W = max(args.warmup, 3) + max(args.settle, 0)       # untimed ticks: settle + warm-up
    KL = max(K, args.latency_ticks)                     # ticks of the latency window (its first K ticks are the throughput window's)
    T = W + KL
    eng = mpcqp.Engine(batch=B, device=local_rank, mode=args.mode)
    peaks = mpcqp.measure_fp64_peak(local_rank)

    # ---- untimed: closed loop through the engine to produce the input sequence (pinned host copies)
    sc = Scenario(B, gaits="trot", seed=20260 + rank)
    h_x = torch.empty((T, B, 12, N + 1), dtype=torch.float64, pin_memory=True)
    h_f = torch.empty((T, B, 20, 13), dtype=torch.float64, pin_memory=True)
    hx, hf = h_x.numpy(), h_f.numpy()
    tot_sweeps = tot_iters = tot_fb = 0.0
    unsolved = 0
    for t in range(T):
        xr, fs = sc.inputs()
        hx[t], hf[t] = xr, fs
        eng.run(t, hx[t], hf[t])
        x = eng.solution()
        info = eng.info(with_y=False)
        if W <= t < W + K:
            tot_sweeps += float(info["sweeps"].sum()); tot_iters += float(info["iters"].sum())
            tot_fb += float((info["iters"] > 0).sum())
        unsolved += int((info["status"] != 1).sum())
        sc.advance(x[:, :12] + xr[:, :, 1])
    d_x, d_f = h_x.cuda(non_blocking=False), h_f.cuda(non_blocking=False)
    esz = 8
    # bytes that cross PCIe per tick: xref in full; of each 20 x 13 gait table only rows 0..7 when row 7 is a terminator for the
    # whole batch (mpcqp_api.cu: copy_gait_tables), which holds for every table of this workload
    brief = bool((hf[:, :, 7, 0] == 0.0).all())
    in_bytes = B * (12 * (N + 1) + (8 * 13 if brief else 260)) * esz

    # ---- timed, device resident: exactly K ticks between barrier + synchronise
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))

    def device_window(n_ticks, overlap):
        """overlap == 1: one tick at a time on the engine's stream, CUDA events around every tick (per-tick latency).
        overlap >= 2: the ticks are issued as that many independent index ranges (mpcqp_set_overlap: a robot's tick t + 1 is ordered
        behind its own tick t only, so the last sweeps of a tick overlap the next tick of the other ranges); two events around the
        whole window, the second one behind mpcqp_join."""
        eng.set_overlap(overlap)
        eng.reset_warm_start()
        for t in range(W):
            eng.run_device(t, d_x[t].data_ptr(), d_f[t].data_ptr())
        a = [torch.cuda.Event(enable_timing=True) for _ in range(n_ticks)]
        b = [torch.cuda.Event(enable_timing=True) for _ in range(n_ticks)]
        eng.synchronize()
        barrier()
        l0 = eng.launches
        if overlap > 1:
            a[0].record(stream)
            for i in range(n_ticks):
                eng.run_device(W + i, d_x[W + i].data_ptr(), d_f[W + i].data_ptr())
            eng.join()
            b[n_ticks - 1].record(stream)
            eng.synchronize()
            barrier()
            eng.set_overlap(1)
            return a[0].elapsed_time(b[n_ticks - 1]), None, eng.launches - l0
        for i in range(n_ticks):
            a[i].record(stream)
            eng.run_device(W + i, d_x[W + i].data_ptr(), d_f[W + i].data_ptr())
            b[i].record(stream)
        eng.synchronize()
        barrier()
        return a[0].elapsed_time(b[n_ticks - 1]), np.array([x.elapsed_time(y) for x, y in zip(a, b)]), eng.launches - l0

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.6)                                     # let nvidia-smi finish starting up (it takes driver locks while it does)
    stagewise_mode = bool(eng.params.mode & 4) and (not (eng.params.mode & 2) or bool(eng.params.mode & 8))
    overlap = args.overlap if args.overlap > 0 else ((8 if world <= 2 else 4) if B <= 6144 else 2)     # 16 launches per tick and process: with 4+ processes on one host 4 ranges are faster
    if not stagewise_mode:
        overlap = 1                                      # the dense path has no index ranges
    total_ms, _, launches = device_window(K, overlap)
    total_ms = max_over_ranks(total_ms)
    value = world * B * K / (total_ms * 1e-3)
    lat_ms, lat_steps, _ = device_window(KL, 1)          # one tick at a time: per-tick latency; its first K ticks are the timed ones
    per_step = lat_steps[:K]
    serial_ms = max_over_ranks(float(per_step.sum()))

    # ---- timed, end to end through the host API (pinned host in, forces out)
    h_out = torch.empty((B, 12), dtype=torch.float64, pin_memory=True)
    out_np = h_out.numpy()

    def host_window(n_ticks):
        eng.reset_warm_start()
        for t in range(W):
            eng.run(t, hx[t], hf[t])
            eng.forces(out=out_np)
        barrier()
        lat = []
        w0 = time.perf_counter()
        for i in range(n_ticks):
            s0 = time.perf_counter()
            eng.run(W + i, hx[W + i], hf[W + i])
            eng.forces(out=out_np)                      # D2H of the step's result, synchronises
            lat.append(time.perf_counter() - s0)
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - w0
        barrier()
        return dt_s, np.array(lat)

    e2e_s, lat = host_window(K)
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * B * K / e2e_s
    checksum = float(np.abs(out_np).sum())
    forces_k = out_np.copy()
    e2e_lat_s, e2e_lat = host_window(KL) if KL > K else (e2e_s, lat)

    # ---- the same K ticks through the asynchronous result protocol (MPC_Wrapper.py:116-260, the control loop picks up
    #      the forces of the previous solve while the current one runs): tick i + 1 is issued before the forces of tick i are waited for
    def async_window(n_ticks):
        eng.set_overlap(min(overlap, 4))                # host inputs: every index range stages its own rows on its own stream (6 calls per
                                                        # range and tick: more than 4 ranges make this leg host-bound; measured 14.9 / 17.5 / 17.4 / 14.6 M at 2 / 3 / 4 / 8)
        eng.reset_warm_start()
        for t in range(W):
            eng.run(t, hx[t], hf[t])
            eng.forces(out=out_np)
        barrier()
        w0 = time.perf_counter()
        for i in range(n_ticks):
            eng.run(W + i, hx[W + i], hf[W + i])
            eng.result_async(i & 1)
            if i > 0:
                eng.result_wait((i - 1) & 1, out_np)
        eng.result_wait((n_ticks - 1) & 1, out_np)
        dt_s = time.perf_counter() - w0
        barrier()
        eng.set_overlap(1)
        return dt_s

    async_s = max_over_ranks(async_window(K))
    async_ok = bool(np.abs(out_np - forces_k).max() <= 1e-9)

    # ---- the same K ticks through the reference-facing classes: MPC_Wrapper.solve(k, planner) + get_latest_result()
    #      (MPC_Wrapper.py:39-78, processing.py:142-145), host arrays in, forces out
    dropin = None
    if not args.no_dropin:
        import MPC_Wrapper

        class _Planner:                                 # what processing.process_mpc hands over: .xref and .fsteps
            pass

        wr = MPC_Wrapper.MPC_Wrapper(0.02, N, 20, 0.32, multiprocessing=False, device=local_rank, mode=args.mode)
        pl = _Planner()
        for t in range(W):
            pl.xref, pl.fsteps = hx[t], hf[t]
            wr.solve(20 * t, pl)
            f_d = wr.get_latest_result()
        barrier()
        dl = []
        w0 = time.perf_counter()
        for i in range(K):
            s0 = time.perf_counter()
            pl.xref, pl.fsteps = hx[W + i], hf[W + i]
            wr.solve(20 * (W + i), pl)
            f_d = wr.get_latest_result()
            dl.append(time.perf_counter() - s0)
        d_s = max_over_ranks(time.perf_counter() - w0)
        barrier()
        dropin = {"value": world * B * K / d_s, "unit": UNIT, "ms_per_step": 1e3 * d_s / K,
                  "latency_ms_p50": 1e3 * float(np.percentile(dl, 50)), "latency_ms_p99": 1e3 * float(np.percentile(dl, 99)),
                  "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 2 * B * 12 * esz,
                  "forces_match_engine_path": bool(np.abs(np.asarray(f_d) - forces_k).max() <= 1e-9),
                  "what": "MPC_Wrapper.solve(k, planner) + get_latest_result() per tick (the drop-in classes of mpc-tsid_b200/), "
                          "forces + first predicted state read back"}
        wr.mpc._engine.close()
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel.  Neither HBM nor tensor bound: the stage-wise kernel is a chain of
    #      6x6 FP64 factorisations per robot, bounded by FP64-FMA-pipe latency/throughput; the denominator is the FP64
    #      FMA peak measured on this GPU in this run.  The fallback (dense ADMM) kernel's work is counted with its own model.
    stagewise = bool(eng.params.mode & 4)
    if stagewise:
        flop = (tot_sweeps * flops_stagewise_sweep() + B * K * flops_stagewise_fixed()
                + tot_iters * flops_stagewise_sweep())            # an interior-point iteration is one stage-wise factorisation + solve
        peak, peak_name = peaks["dfma_tflops"], "FP64 FMA"
        kernel = "riccati_kernel<16, true> (+ ipm_kernel<16> for the robots the sweeps give up on)"
        prof = "profiles/r02_riccati_kernel_ncu_summary.txt"
        if not os.path.exists(os.path.join(ROOT, prof)):
            prof = "profiles/r01_riccati_kernel_ncu_summary.txt"
        model = "engine's own count, DESIGN.md section 5: sweeps*%.0f + %.0f per solve (+ fallback work)" % (
            flops_stagewise_sweep(), flops_stagewise_fixed())
    else:
        flop = flops_per_tick(tot_sweeps, tot_iters, tot_fb)        # this rank, K ticks
        peak, peak_name = peaks["dmma_tflops"], "FP64 tensor (DMMA m8n8k4)"
        kernel = "solve_kernel<16,false> (+ ADMM fallback kernel)"
        prof = "profiles/r01_solve_kernel_ncu_summary.txt"
        model = "engine's own count, DESIGN.md section 5: sweeps*%.0f + admm_iters*%.0f per solve" % (
            flops_per_tick(1, 0, 0) - flops_per_tick(0, 0, 0), flops_per_tick(0, 1, 0) - flops_per_tick(0, 0, 0))
    # DRAM traffic of the dominant kernel: read from the committed ncu summary of one 4096-robot launch (not measured in this run:
    # bench.py never runs under a profiler); null when the batch differs from the profiled one
    rd, wrb = profile_metric(prof, "dram__bytes_read.sum"), profile_metric(prof, "dram__bytes_write.sum")
    traffic = (rd + wrb) if (rd is not None and wrb is not None and B == 4096) else None
    traffic_src = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at 4096 robots, parsed at run time from "
                   "the committed ncu --set full summary %s" % prof)
    pipe_busy = profile_metric(prof, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    flop_all = sum_over_ranks(flop)
    achieved = flop_all / world / (total_ms * 1e-3) * 1e-12          # per GPU
    hbm_alg = 21216.0                                               # B per solve, SURVEY.md 8(d)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "measured"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback"
    hbm_ach = hbm_alg * B * K / (total_ms * 1e-3) * 1e-9
    canon = 4.31e6 * world * B * K / (total_ms * 1e-3) * 1e-12 / world
    roofline = {
        "bound": "fp64", "bound_note": "neither hbm nor tensor: FP64 FMA pipe latency/throughput (SURVEY.md 8d)", "pipe": peak_name,
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "traffic": traffic, "kernel": kernel, "traffic_source": traffic_src,
        "frac_note": "frac counts USEFUL FP64 work (one copy of every 6x6 factorisation per robot); the kernel factors redundantly in "
                     "every lane, so the FP64 pipe is busier than that: executed_fp64_pipe_busy_frac (ncu, same committed profile)",
        "executed_fp64_pipe_busy_frac": None if pipe_busy is None else pipe_busy / 100.0,
        "peak_source": "issue loop of that instruction measured on this GPU in this run (mpcqp_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "flop_model": model,
        "mflop_per_solve": flop / (B * K) * 1e-6, "canonical_mflop_per_solve_survey_8d_it60": 4.31,
        "canonical_equivalent_tflops": canon,
        "sweeps_per_solve": tot_sweeps / (B * K), "fallback_iters_per_solve": tot_iters / (B * K),
        "fallback_frac": tot_fb / (B * K),
        "dfma_peak_tflops": peaks["dfma_tflops"], "dmma_peak_tflops": peaks["dmma_tflops"],
        "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                "bytes_per_solve": hbm_alg, "peak_source": hbm_src},
    }

    eng.close()
    del d_x, d_f
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, short device-resident sub-runs (planner + QP + integration per tick on the GPU)
    other = None
    if not args.no_other_configs:
        sw = lambda **kw: run_sweep(torch, mpcqp, Scenario, rank, local_rank, world, barrier, max_over_ranks, W=25, e2e=False,
                                    peak_tflops=peaks["dfma_tflops"], **kw)
        other = {
            "configs[2] mixed gaits N=16, 65536 robots per GPU": sw(B=65536, K=20, gaits=["trot", "pace", "bound", "walk"]),
            "configs[3] long horizon N=32, 4096 trot robots per GPU": sw(B=4096, K=20, gaits=["trot"], n_steps=32),
            "configs[3] long horizon N=64, 2048 trot robots per GPU": sw(B=2048, K=20, gaits=["trot"], n_steps=64),
            "configs[4] closed-loop sweep N=16, 131072 trot robots per GPU": sw(B=131072, K=20, gaits=["trot"]),
            "what": "mpcqp_scenario_run: footstep planner + QP build + solve + centroidal state integration per tick inside the solve "
                    "kernel, 25 untimed closed-loop ticks first; value = robots x ticks / device time (max over ranks), whole job",
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                 # the CPU arm gets every host core back
        cpu, _ = cpu_arm(args.cpu_ticks, 3)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % B,
                       "instances_per_gpu": B, "parallelism": "instances sharded by index, no collective",
                       "cpu_affinity": affinity,
                       "settle_ticks": max(args.settle, 0),
                       "tick_overlap": ("the K timed ticks are issued back to back as %d independent index ranges of the batch "
                                        "(mpcqp_set_overlap): a robot's tick t+1 runs behind its own tick t only (the warm start, MPC.py:403-406), so "
                                        "the last sweeps of one tick overlap the next tick of the other ranges; results are bit-identical to "
                                        "one tick at a time (value_one_tick_at_a_time, latency_ms)" % overlap) if overlap > 1 else "off",
                       "tick_window": "ticks %d..%d of the closed loop are timed (steady operation; the first %d ticks after "
                                      "release from rest run untimed before the %d warm-up ticks)" % (W, W + K - 1, max(args.settle, 0), max(args.warmup, 3)),
                       "l2": "each timed tick reads its own input block (%d x %.1f MB > 126 MB L2 over the run); the "
                             "carried warm-start state (%.1f MB) is hot by design" % (K, B * (12 * (N + 1) + 260) * esz / 1e6, B * 4.2e-3)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": B * 12 * esz,
                    "host_buffer_bytes_per_step": B * (12 * (N + 1) + 260) * esz,
                    "ms_per_step": 1e3 * e2e_s / K, "latency_ms_p50": 1e3 * float(np.percentile(lat, 50)),
                    "latency_ms_p99": 1e3 * float(np.percentile(lat, 99)),
                    "what": "mpcqp_run(host xref, host fsteps) + mpcqp_get_latest_result(host forces) per tick through the C ABI; the inputs sit in "
                            "page-locked host memory and are fetched over PCIe by the solve kernel itself, robot by robot (no staging copy); "
                            "h2d_bytes_per_step = the bytes that cross the bus"},
            "e2e_async": {"value": world * B * K / async_s, "unit": UNIT, "ms_per_step": 1e3 * async_s / K,
                          "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": B * 12 * esz, "forces_match_synchronous_path": async_ok,
                          "what": "mpcqp_run(host inputs of tick i + 1) is issued before mpcqp_result_wait(forces of tick i): the asynchronous "
                                  "protocol of MPC_Wrapper.py:116-260 (results one tick late), the ticks issued as overlapped index ranges that stage "
                                  "their own inputs and copy their own forces back; not the headline e2e"},
            "e2e_dropin": dropin,
            "gpu_launches": int(launches),
            "value_one_tick_at_a_time": world * B * K / (serial_ms * 1e-3),
            "latency_ms": {"p50": float(np.percentile(per_step, 50)), "p99": float(np.percentile(per_step, 99)),
                           "what": "device time of one batch tick (CUDA events on the engine stream), the K timed ticks"},
            "latency_window": {"ticks": int(KL), "after_untimed_ticks": int(W),
                               "device_ms_p50": float(np.percentile(lat_steps, 50)), "device_ms_p99": float(np.percentile(lat_steps, 99)),
                               "device_ms_max": float(lat_steps.max()),
                               "e2e_ms_p50": 1e3 * float(np.percentile(e2e_lat, 50)), "e2e_ms_p99": 1e3 * float(np.percentile(e2e_lat, 99)),
                               "e2e_ms_max": 1e3 * float(e2e_lat.max()),
                               "what": "SURVEY.md 8d: one batch tick from enqueue to forces readable, p50 / p99 over this many consecutive "
                                       "closed-loop ticks (a separate pass over the same inputs; its first %d ticks are the timed ones)" % K},
            "roofline": roofline,
            "other_configs": other,
            "cpu_baseline": cpu,
            "reference_arm_cores": None if cpu is None else cpu["cores"],
            "clocks": clocks,
            "unsolved_instances": unsolved, "forces_checksum": checksum,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
