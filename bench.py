#!/usr/bin/env python
"""Headline benchmark: full gradient-descent iterations of the 10 000-buoy square-mesh OCP (BASELINE.json cfg3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One *step* = one gradient-descent iteration of Pipeline_limits.py at its defaults (f = (0.1, 0), no line search):
forward Navier-Stokes (Newton, 3 its), grad(u) projection, primal buoy ODE, backward sweep (adjoint ODE + point
sources + misfit), [all-reduce], adjoint Navier-Stokes, gradient, control update, cost.  Every step restarts from
q0 so that each step is exactly iteration 0 of the reference run (with LR = 5 the reference itself diverges
after four iterations).  --gpus N shards THE SAME 10 000 buoys over the N ranks (BASELINE config 3: strong scaling,
K_global = 10 000, alpha = 1e-6 K_global, one NCCL all-reduce of [b | misfit | n_masked] per step inside the C
library); the replicated FE solve bounds that curve (Amdahl), which the line states.  The north-star scaling
target - the 1e7-buoy synthetic sweep on the refined 128 x 128 mesh - is measured in the same run and reported as
the `sweep_strong` object, with the single-GPU time of the same sweep taken in-line on rank 0.

metric  = buoy-steps/s through full GD iterations = 3 sweeps x K_global x 200 samples per iteration / time
          (the published 1500 s/iteration = 6.7e-4 it/s = 4.0e3 buoy-steps/s).
value   = inputs resident in HBM; e2e = control / results cross the host boundary every step.
roofline: the backward sweep (adjoint ODE + scatter + misfit), HBM-bound by design: 48 B per buoy-step;
roofline_lu: the sparse factorisation (flops of the symbolic analysis / measured fp64 FMA peak) and one
triangular-solve pass (bytes of L+U / measured HBM peak) - the kernels that dominate the step by time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_GLOBAL = 10_000
NT = 200
PUBLISHED_BUOY_STEPS_PER_S = 4.0e3          # BASELINE.md section 1 (derived from 1500 s / iteration)
# dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full captures (profiles/)
NCU_TRAFFIC_INSTEP_BACKWARD = 114_538_752 + 4_786_944         # K = 10 000, time-parallel sweep (two passes over the streams,
                                                              # the second mostly from L2): algorithmic 96 000 000 B
NCU_TRAFFIC_SWEEP_BACKWARD = 10_081_863_000 + 5_742_000       # K = 2^20:   algorithmic 10 066 329 600 B
NCU_TRAFFIC_SWEEP_FORWARD = 19_127_000 + 6_656_207_000        # K = 2^20:   algorithmic  6 710 886 400 B
METRIC = "gd_buoy_steps_per_sec"
UNIT = "buoy-steps/s (3 sweeps x K x 200 per GD iteration)"


def reference_grid():
    """cfg3 start points: 100 x 100 grid on [0.1,0.4] x [0.25,1.75], x fastest."""
    gx, gy = np.meshgrid(np.linspace(0.1, 0.4, 100), np.linspace(0.25, 1.75, 100))
    return np.stack([gx.ravel(), gy.ravel()], 1)


def golden_field():
    f = np.load(os.path.join(ROOT, "tests", "golden", "fields.npz"))
    return f["velocity_100"]                # identical to reference_runs/10000_buoys/paraview/velocity.h5


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_gd_iteration(P, f0, LR):
    """One GD iteration with the oracle (CPU restatement of the reference's path)."""
    s = P.gradient_step(f0)
    f1 = f0 - LR * s["grad"]
    return P.cost(s["u"], f1)


def make_cpu_pipeline(K, threads):
    import ocp_b200  # noqa: F401
    from ocp_b200.fespace import TaylorHood
    from ocp_b200.mesh import square_mesh
    from ocp_b200.pipeline import initial_control
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    V = TaylorHood(square_mesh(32))
    x0 = reference_grid()[:K]
    from oracle.buoy_oracle import BuoyOracle
    B = BuoyOracle(V)
    B.set_threads(threads)
    _, ud, *_ = B.forward(V.velocity_nodal(golden_field()), x0, NT, 0.005, [1.0, 1.0])
    P = helpers.OraclePipeline(V, 1.0, x0, ud, 1e-6 * K)
    P.B.set_threads(threads)
    return P, initial_control(V, "PL")


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU algorithm for the same step on the host cores.  FEniCS is not
    installable here, so this is the oracle port: per-buoy loops in C over all host threads, FE assembly in NumPy,
    SuperLU (SciPy) for the solves - the same workload (10 000 buoys, full GD iteration), same K_global for every
    --gpus N (the GPU arm shards these very buoys)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    P, f0 = make_cpu_pipeline(K_GLOBAL, threads)
    warm = max(0, min(args.warmup, 5))
    for _ in range(warm):
        cpu_gd_iteration(P, f0, 5.0)
    steps = max(1, min(args.steps, 50))          # ~0.8 s per step: K = 10 stays well inside "a few minutes"
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_gd_iteration(P, f0, 5.0)
    dt = (time.perf_counter() - t0) / steps
    val = 3 * K_GLOBAL * NT / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": val / PUBLISHED_BUOY_STEPS_PER_S, "dtype": "f64", "data": "synthetic",
        "gd_iters_per_sec": 1.0 / dt,
        "config": cfg3_config(1, K_GLOBAL, None),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} full GD iterations of the whole 10000-buoy workload (no sub-sampling); "
                                   "buoy loops on all host threads, FE solves single-threaded SuperLU"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cfg3_config(world, K_local, ocp):
    cfg = {"workload": "cfg3 square N=32 OCP, 10000 buoys (the reference's 100x100 grid), Pipeline_limits defaults, "
                       "GD iteration 0",
           "K_global": K_GLOBAL, "nt": NT, "mesh": "square 32x32", "ndofs": 9539}
    if ocp is not None:
        cfg.update({"K_per_gpu": K_local, "nnz": int(ocp.V.csr_col.size), "newton_its": ocp.last_newton_its,
                    "parallelism": f"the 10000 buoys sharded x{world}, replicated FE solve, 1 ncclAllReduce/step "
                                   "inside libocp_b200 (ocp_allreduce)",
                    "l2": "256 MiB buffer written between timed iterations (L2 flush)"})
    return cfg


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def quiet_init(backend="nccl"):
    """Process group from torchrun's environment.  NCCL may print its banner on stdout; the contract is ONE JSON line
    there, so fd 1 is routed to stderr while the communicator is created (first collective included)."""
    import torch
    import torch.distributed as dist
    from ocp_b200.sharding import init_from_env
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        group, rank, world, local = init_from_env(backend)
        if group is not None:
            t_ = torch.zeros(1, device=torch.device("cuda", local))
            dist.all_reduce(t_, group=group)
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    return group, rank, world, local


def sweep_measure(V, d_w_host, Kt, lo, hi, dev, group, steps, warm):
    """Forward + backward sweep (+ all-reduce) of buoys [lo, hi) of the Kt-buoy synthetic set (cfg5: uniform start
    points, numpy default_rng(0); u_d = 1.1 x the twin velocities).  Returns (ms per step, misfit, ||b||)."""
    import torch
    import torch.distributed as dist
    from ocp_b200.pipeline import OCP, Parameters
    rng = np.random.default_rng(0)
    x0 = np.stack([rng.uniform(0.1, 1.9, Kt), rng.uniform(0.1, 1.9, Kt)], 1)[lo:hi]
    ocp = OCP(V, Parameters(), x0, None, device=dev, group=group, alpha_scale_K=Kt)
    del x0
    K = hi - lo
    d_w = torch.from_numpy(d_w_host).to(dev)
    ocp.ctx.project_grad(d_w, ocp.d_g)
    ocp._primal(d_w, ocp.d_x, ocp.d_u, ocp.d_mask)
    torch.mul(ocp.d_u, 1.1, out=ocp.d_ud)

    def step():
        ocp._primal(d_w, ocp.d_x, ocp.d_u, ocp.d_mask)
        ocp.d_acc.zero_()
        ocp.ctx.buoy_adjoint_scatter(ocp.d_vel, ocp.d_g, K, ocp.d_x, ocp.d_u, ocp.d_ud, ocp.d_mask, ocp.d_parked,
                                     None, ocp.d_acc)
        ocp._allreduce(ocp.d_acc)

    for _ in range(warm):
        step()
    if group is not None:
        dist.barrier(group=group)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if group is not None:
        dist.barrier(group=group)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if group is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    nn = V.num_nodes
    out = (float(ms.item()) / steps, float(ocp.d_acc[2 * nn].item()), float(ocp.d_acc[:2 * nn].norm().item()))
    ocp.close()
    del ocp
    torch.cuda.empty_cache()
    return out


def sweep_strong(args, dev, group, rank, world, peak):
    """The north-star scaling measurement inside the default run: K_total = 1e7 synthetic buoys on the refined
    128 x 128 mesh, sharded over the ranks (strong scaling), against the single-GPU time of the SAME sweep taken
    in-line on rank 0 (the other ranks wait)."""
    import torch
    import torch.distributed as dist
    from ocp_b200.fespace import TaylorHood
    from ocp_b200.mesh import square_mesh
    from ocp_b200.pipeline import OCP, Parameters, initial_control
    from ocp_b200.sharding import shard_bounds
    Kt, n = args.sweep_buoys, args.sweep_mesh
    V = TaylorHood(square_mesh(n))
    # field: forward solve with f = (0.1, 0) on that mesh (every rank solves it redundantly: it is the replica)
    tmp = OCP(V, Parameters(), np.array([[1.0, 1.0]]), None, device=dev)
    w_host = tmp.forward_solve(torch.from_numpy(initial_control(V, "PL")).to(dev)).d_w.cpu().numpy()
    tmp.close()
    del tmp
    lo, hi = shard_bounds(Kt, rank, world)
    steps, warm = max(2, min(args.steps, 5)), 3
    ms_n, misfit_n, bnorm_n = sweep_measure(V, w_host, Kt, lo, hi, dev, group, steps, warm)
    ms_1 = misfit_1 = bnorm_1 = None
    if world == 1:
        ms_1, misfit_1, bnorm_1 = ms_n, misfit_n, bnorm_n
    else:
        if rank == 0:
            ms_1, misfit_1, bnorm_1 = sweep_measure(V, w_host, Kt, 0, Kt, dev, None, steps, warm)
        dist.barrier(group=group)
    if rank != 0:
        return None
    units = 2.0 * Kt * NT
    gbs = 80.0 * (hi - lo) * NT / (ms_n * 1e-3) / 1e9
    return {"workload": f"cfg5 synthetic sweep: {Kt} buoys total on square {n}x{n}, forward + backward sweep + "
                        "all-reduce per step, sharded over the ranks (strong scaling)",
            "K_total": Kt, "K_per_gpu": hi - lo, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_n, "buoy_steps_per_sec": units / (ms_n * 1e-3),
            "single_gpu_ms_per_step_inline": ms_1, "single_gpu_buoy_steps_per_sec": units / (ms_1 * 1e-3),
            "speedup_vs_1gpu": ms_1 / ms_n, "target_speedup_at_8": 7.0,
            "per_gpu_hbm_gbs": gbs, "per_gpu_frac_of_hbm_peak": gbs / peak,
            "parity_vs_single_gpu": {"misfit_rel_diff": abs(misfit_n - misfit_1) / abs(misfit_1),
                                     "b_norm_rel_diff": abs(bnorm_n - bnorm_1) / abs(bnorm_1)}}


# ------------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ocp_b200  # noqa: F401
    from ocp_b200 import capi
    from ocp_b200.fespace import TaylorHood
    from ocp_b200.mesh import square_mesh
    from ocp_b200.pipeline import OCP, Parameters, State, initial_control
    from ocp_b200.sharding import shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    group, rank, world, local = quiet_init("nccl")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    V = TaylorHood(square_mesh(32))
    P = Parameters()
    lo, hi = shard_bounds(K_GLOBAL, rank, world)
    x0_all = reference_grid()
    x0 = np.ascontiguousarray(x0_all[lo:hi])
    K = x0.shape[0]
    f0 = initial_control(V, "PL")
    LR = 5.0
    # u_d regenerated from the stored 10000_buoys field with the primal kernel itself
    ocp = OCP(V, P, x0, np.zeros((K, NT, 2)), device=dev, group=group)
    assert ocp.K_global == K_GLOBAL
    d_field = torch.from_numpy(golden_field()).to(dev)
    ocp._primal(d_field, ocp.d_x, ocp.d_u, ocp.d_mask)
    ocp.d_ud.copy_(ocp.d_u)
    ud_host = ocp._to_reference_layout(ocp.d_ud)
    ctx = ocp.ctx
    ctx.set_observations_host(x0, ud_host)
    nn = V.num_nodes
    d_f0 = torch.from_numpy(f0).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)     # > 126 MB L2

    def barrier():
        if group is not None:
            dist.barrier(group=group)
        torch.cuda.synchronize()

    def step_resident():
        flush.zero_()
        ocp.d_f.copy_(d_f0)
        ocp.gradient_step(ocp.d_f)
        ocp.ctx.nodal_axpby(1.0, ocp.d_f, -LR, ocp.d_grad, ocp.d_f)       # f <- f - LR (alpha f - z), OCP_dolfin.py:426
        return ocp._cost_from_acc(ocp.d_f)

    h_f = torch.from_numpy(f0.copy()).pin_memory()
    h_grad = torch.empty((nn, 2), dtype=torch.float64).pin_memory()
    h_sc = torch.empty(4, dtype=torch.float64).pin_memory()

    def step_e2e():
        """Host control in, host gradient / cost / mask count out, through the public Python API."""
        flush.zero_()
        ocp.d_f.copy_(h_f, non_blocking=True)                     # H2D: control
        ocp.gradient_step(ocp.d_f)
        ocp.ctx.boundary_inner(ocp.d_f, ocp.d_f, ocp.d_sc)
        h_grad.copy_(ocp.d_grad, non_blocking=True)               # D2H: gradient field
        h_sc[:2].copy_(ocp.d_acc[2 * nn:], non_blocking=True)     # D2H: misfit, n_masked
        h_sc[2:3].copy_(ocp.d_sc[:1], non_blocking=True)
        torch.cuda.synchronize()
        return float(h_sc[0] + 0.5 * ocp.alpha * h_sc[2])     # J(u, f) at the incoming control

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if group is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
        return float(ms.item()) / steps, out

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(30):            # set-up, not measurement: bring the SM clocks up before the W warm-up steps
        step_resident()
    torch.cuda.synchronize()
    n_launch0 = capi.launch_count()
    # `ncu --profile-from-start off ... python bench.py` lists exactly the launches of the timed steps
    prof_region = os.environ.get("OCP_BENCH_PROFILE_REGION") == "1"
    if prof_region:
        for _ in range(warm):
            step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    ms_step, J = timed(step_resident, args.steps, warm)
    launches = (capi.launch_count() - n_launch0) / (args.steps + warm)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, J2 = timed(step_e2e, args.steps, 3)
    # the C-ABI host-buffer call, collective included (ocp_gradient_host all-reduces inside the library)
    out = (np.empty(V.ndofs), np.empty(V.ndofs), np.zeros(K), np.zeros(4))
    ms_cabi, _ = timed(lambda: (flush.zero_(), ctx.gradient_host(f0, out))[1], args.steps, 3)
    # line items: a separate pass with per-phase event timing switched on (it synchronises after every phase,
    # so it is kept out of the timed regions above)
    nprof = 3
    ctx.set_profiling(True)
    step_resident()
    ctx.reset_solver_stats()
    for _ in range(nprof):
        step_resident()
    stats = ctx.solver_stats()
    ctx.set_profiling(False)
    info = ctx.solver_info()

    # ---- the hand-written buoy kernels, timed live with CUDA events
    def back():
        ocp.d_acc.zero_()
        ctx.buoy_adjoint_scatter(ocp.d_vel, ocp.d_g, K, ocp.d_x, ocp.d_u, ocp.d_ud, ocp.d_mask, ocp.d_parked, None,
                                 ocp.d_acc)

    def kernel_ms(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps

    ocp.d_f.copy_(d_f0)
    ocp.gradient_step(ocp.d_f)        # leave the context in its in-step state (w, g, trajectories of the 10000 buoys)
    ms_back = kernel_ms(back)
    ms_fwd = kernel_ms(lambda: ocp._primal(ocp.d_w, ocp.d_x, ocp.d_u, ocp.d_mask))
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    bytes_back = 48.0 * K * NT
    ach = bytes_back / (ms_back * 1e-3) / 1e9

    # ---- sparse LU roofline: flops of the symbolic analysis / measured fp64 FMA peak; one solve pass streams L+U once
    fp64_peak = ctx.fp64_peak_tflops()
    lu = None
    if stats["n_factor"] > 0 and stats["n_solve"] > 0:
        ms_factor = stats["factor_ms"] / stats["n_factor"]
        ms_solve = stats["solve_ms"] / stats["n_solve"]
        fl = info["factor_flops"]
        by = 8.0 * info["factor_nnz"]
        lu = {"factor": {"bound": "fp64", "flops_per_launch_sequence": fl, "ms": ms_factor,
                         "achieved": fl / (ms_factor * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": fl / (ms_factor * 1e-3) / 1e12 / fp64_peak,
                         "peak_source": "measured live: register-resident DFMA loop on all SMs (ocp_selftest_fp64_peak)"},
              "solve_pass": {"bound": "hbm", "bytes_per_pass": by, "ms": ms_solve,
                             "achieved": by / (ms_solve * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": by / (ms_solve * 1e-3) / 1e9 / peak,
                             "note": "the factors are L2-resident at this size; a pass is a dependent chain of "
                                     "pivot blocks per front and tree level (latency-bound)"},
              "levels": info["levels"], "fronts": info["fronts"], "max_front": info["max_front"]}

    # ---- the same buoy kernels at sweep size (2^20 buoys on the same mesh) where they are bandwidth-relevant
    sweep = None
    if rank == 0 and not args.no_sweep:
        Ks = 1 << 20
        rng = np.random.default_rng(0)
        xs = np.stack([rng.uniform(0.1, 1.9, Ks), rng.uniform(0.1, 1.9, Ks)], 1)
        big = OCP(V, P, xs, None, device=dev)
        big.ctx.project_grad(ocp.d_w, big.d_g)
        big._primal(ocp.d_w, big.d_x, big.d_u, big.d_mask)
        torch.mul(big.d_u, 1.1, out=big.d_ud)

        def bback():
            big.d_acc.zero_()
            big.ctx.buoy_adjoint_scatter(big.d_vel, big.d_g, Ks, big.d_x, big.d_u, big.d_ud, big.d_mask,
                                         big.d_parked, None, big.d_acc)
        mb = kernel_ms(bback, 5)
        mf = kernel_ms(lambda: big._primal(ocp.d_w, big.d_x, big.d_u, big.d_mask), 5)
        sweep = {"K": Ks, "backward_ms": mb, "forward_ms": mf,
                 "backward_gbs": 48.0 * Ks * NT / (mb * 1e-3) / 1e9, "forward_gbs": 32.0 * Ks * NT / (mf * 1e-3) / 1e9,
                 "backward_frac_of_peak": 48.0 * Ks * NT / (mb * 1e-3) / 1e9 / peak,
                 "forward_frac_of_peak": 32.0 * Ks * NT / (mf * 1e-3) / 1e9 / peak,
                 "buoy_steps_per_sec_fwd_plus_bwd": 2.0 * Ks * NT / ((mb + mf) * 1e-3),
                 "backward_traffic": NCU_TRAFFIC_SWEEP_BACKWARD, "forward_traffic": NCU_TRAFFIC_SWEEP_FORWARD,
                 "traffic_source": "profiles/prof_buoy_final.raw.txt"}
        big.close()
        del big
        torch.cuda.empty_cache()

    # ---- per-function host boundary: the reference's own call shapes with (K,200,2) host arrays
    per_fn = None
    if rank == 0 and not args.no_sweep:
        w_host = ocp.d_w.cpu().numpy()
        g_host = ocp.d_g.cpu().numpy()
        mask = np.zeros(K)
        xh, uh = ctx.solve_primal_ode_host(w_host, x0, mask)

        def wall(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3
        per_fn = {
            "forward_solve_host_f_in_w_out": wall(lambda: ocp.forward_solve(f0).vector()),
            "solve_primal_ode_host": wall(lambda: ctx.solve_primal_ode_host(w_host, x0, mask)),
            "solve_adjoint_ode_host": wall(lambda: ctx.solve_adjoint_ode_host(g_host, xh, uh, ud_host, mask)),
            "bytes_host_arrays_per_call": int(xh.nbytes),
            "note": "wall clock, pageable numpy arrays as the reference holds them: x, u, u_d, mu are (K,200,2) f8 = "
                    "32 MB each at K = 10000, so these calls are PCIe/pageable-copy bound; the one-call API "
                    "(ocp_gradient_host, e2e above) keeps the trajectories on the device"}
        if world == 1:      # (collective inside: only timed where rank 0 is the whole job)
            per_fn["adjoint_solve_host_x_u_in_z_out"] = wall(
                lambda: ocp.adjoint_solve(State(ocp.d_w), xh, uh, mask, ocp.d_g).vector())

    # ---- sharded vs single-rank parity of the exchanged accumulator (NCCL evidence)
    parity = None
    if world > 1:
        nn2 = 2 * nn
        ocp.d_f.copy_(d_f0)
        ocp.gradient_step(ocp.d_f)
        acc_sh = ocp.d_acc.clone()
        J_sh = ocp._cost_from_acc(ocp.d_f)
        z_sh = ocp.d_z.clone()
        if rank == 0:
            full = OCP(V, P, x0_all, np.zeros((K_GLOBAL, NT, 2)), device=dev)
            full._primal(d_field, full.d_x, full.d_u, full.d_mask)
            full.d_ud.copy_(full.d_u)
            full.d_f.copy_(d_f0)
            full.gradient_step(full.d_f)
            J_1 = full._cost_from_acc(full.d_f)
            b1, bs = full.d_acc[:nn2], acc_sh[:nn2]
            parity = {"J_rel_diff": abs(J_sh - J_1) / abs(J_1),
                      "b_max_rel_diff": float(((bs - b1).abs().max() / b1.abs().max()).item()),
                      "b_norm_rel_diff": float(((bs.norm() - b1.norm()).abs() / b1.norm()).item()),
                      "z_max_rel_diff": float(((z_sh - full.d_z).abs().max() / full.d_z.abs().max()).item()),
                      "misfit_rel_diff": float(((acc_sh[nn2] - full.d_acc[nn2]).abs() / full.d_acc[nn2].abs()).item()),
                      "n_masked_equal": bool(acc_sh[nn2 + 1] == full.d_acc[nn2 + 1]),
                      "what": "all-reduced accumulator of the sharded run (ncclAllReduce inside libocp_b200) vs a "
                              "rank-0 recomputation on the unsharded 10000 buoys"}
            full.close()
            del full
        dist.barrier(group=group)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only): the oracle on the full workload
    cpu = cpu_serial = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        Pc, f0c = make_cpu_pipeline(K_GLOBAL, threads)
        cpu_gd_iteration(Pc, f0c, LR)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            Jc = cpu_gd_iteration(Pc, f0c, LR)
        dtc = (time.perf_counter() - t0) / reps
        cpu = {"value": 3 * K_GLOBAL * NT / dtc, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{reps} full GD iterations of the whole 10000-buoy workload (no sub-sampling); best-effort "
                         "mode: buoy loops on all host threads (OpenMP), FE assembly NumPy, SuperLU solves",
               "ms_per_step": dtc * 1e3, "J": Jc, "J_gpu": J, "J_rel_diff": abs(Jc - J) / abs(Jc)}
        # reference-structured mode: one thread, the reference's loop order (buoy by buoy, sample by sample)
        Pc.B.set_threads(1)
        t0 = time.perf_counter()
        cpu_gd_iteration(Pc, f0c, LR)
        dts = time.perf_counter() - t0
        cpu_serial = {"value": 3 * K_GLOBAL * NT / dts, "unit": UNIT, "cores": 1, "kind": "port",
                      "sample": "1 full GD iteration of the 10000-buoy workload on ONE host thread, per-buoy / "
                                "per-sample loops in the reference's order (the reference itself is serial)",
                      "ms_per_step": dts * 1e3}

    ss = sweep_strong(args, dev, group, rank, world, peak) if not args.no_sweep else None

    if rank == 0:
        units = 3.0 * K_GLOBAL * NT
        val = units / (ms_step * 1e-3)
        items = {"assembly": stats["assemble_ms"] / nprof,
                 "sparse_lu_refactor": stats["factor_ms"] / nprof,
                 "sparse_lu_solve": stats["solve_ms"] / nprof,
                 "buoy_forward_kernel": ms_fwd, "buoy_backward_kernel": ms_back}
        dom = max(items, key=items.get)
        fe_ms = items["assembly"] + items["sparse_lu_refactor"] + items["sparse_lu_solve"]
        line = {
            "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": val / PUBLISHED_BUOY_STEPS_PER_S, "dtype": "f64", "data": "synthetic",
            "gd_iters_per_sec": 1e3 / ms_step,
            "config": cfg3_config(world, K, ocp),
            "e2e": {"value": units / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h_f.numel() * 8), "d2h_bytes_per_step": int(h_grad.numel() * 8 + 24),
                    "api": "OCP.gradient_step via pinned host control / gradient buffers",
                    "cabi_host_call_ms": ms_cabi, "per_function_ms": per_fn},
            "gpu_launches": round(launches, 1),
            "clocks": clocks,
            "line_items_ms_per_step": dict(items, n_factor_per_step=stats["n_factor"] / nprof,
                                           n_solve_per_step=stats["n_solve"] / nprof,
                                           n_dense_operator_applies_per_step=stats["n_dense"] / nprof,
                                           one_time_symbolic_analysis_ms=stats["analyse_ms"]),
            "dominant_by_time": {"item": dom, "ms": items[dom], "share_of_step": items[dom] / ms_step},
            "amdahl": {"replicated_fe_ms": fe_ms, "sharded_buoy_ms": ms_fwd + ms_back,
                       "note": "strong scaling of cfg3 divides only the buoy sweeps; the 9539-dof FE solve is "
                               "replicated on every rank (cheaper than communicating it), so the step is bounded "
                               "below by replicated_fe_ms - see sweep_strong for the scaling target"},
            "roofline": {"bound": "hbm", "kernel": "buoy_adjoint_scatter_tp_kernel (in-step backward sweep)", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": NCU_TRAFFIC_INSTEP_BACKWARD if world == 1 else None,
                         "traffic_source": "profiles/prof_r2_tp.summary.txt (ncu --set full, dram read+write)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_back, "ms": ms_back,
                         "note": "in-step launch (this rank's share of the 10000 buoys): latency-bound, the "
                                 "time-parallel sweep is used at this size; see roofline_sweep for the serial sweep "
                                 "kernels at 2^20 buoys, where the path is bandwidth-bound"},
            "roofline_sweep": sweep,
            "roofline_lu": lu,
            "sweep_strong": ss,
            "parity_vs_single_rank": parity,
            "cpu_baseline": cpu,
            "cpu_baseline_serial": cpu_serial,
            "nccl_version": capi.nccl_version(),
            "J_after_update": J, "J_at_q0_e2e": J2,
        }
        print(json.dumps(line), flush=True)
    ocp.close()
    if group is not None:
        dist.destroy_process_group()


def run_sweep(args):
    """--workload sweep: only the synthetic drifter sweep of BASELINE.json cfg5 (see sweep_strong) as its own line."""
    import torch
    import torch.distributed as dist
    import ocp_b200  # noqa: F401
    group, rank, world, local = quiet_init("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peak = float(load_peaks().get("hbm_gbs", 6650.0))
    ss = sweep_strong(args, dev, group, rank, world, peak)
    if rank == 0:
        print(json.dumps({
            "metric": "sweep_buoy_steps_per_sec", "value": ss["buoy_steps_per_sec"],
            "unit": "buoy-steps/s (forward + backward sweep)", "n_gpus": world, "steps": ss["steps"],
            "warmup": ss["warmup"], "ms_per_step": ss["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ss["workload"], "K_total": ss["K_total"], "K_per_gpu": ss["K_per_gpu"], "nt": NT,
                       "l2": "trajectory arrays (3 x 32 B x K x 200) far exceed L2"},
            "roofline": {"bound": "hbm", "kernel": "buoy_forward + buoy_adjoint_scatter",
                         "achieved": ss["per_gpu_hbm_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": ss["per_gpu_frac_of_hbm_peak"], "traffic": None,
                         "algorithmic_bytes_per_launch": 80.0 * ss["K_per_gpu"] * NT},
            "sweep_strong": ss}), flush=True)
    if group is not None:
        dist.destroy_process_group()


def run_ensemble(args):
    """--workload ensemble: BASELINE.json cfg4 on ONE GPU - initial_control_test.py cases 0-3 (6 buoys) plus the
    Pipeline_limits.py sweep over 10 / 100 / 400 / 10 000 buoys = 8 independent cases, each with its own context,
    CUDA stream and host thread, one gradient evaluation per case per step through the C-ABI host-buffer call.
    Reports aggregate GD iterations/s for the concurrent run and for the same cases run one after the other."""
    import torch
    import ocp_b200  # noqa: F401
    from ocp_b200.ensemble import Case, Ensemble
    from ocp_b200.fespace import TaylorHood
    from ocp_b200.mesh import square_mesh
    from ocp_b200.pipeline import OCP, Parameters, initial_control
    import torch.distributed as dist
    group, rank, world, local = quiet_init("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    V = TaylorHood(square_mesh(32))
    gold = os.path.join(ROOT, "tests", "golden")
    cases = []
    t6 = np.load(os.path.join(gold, "traj_6_buoys.npz"))
    for c in range(4):
        cases.append(Case(f"ICT case {c}", t6["x_0_array"][:, 0, :].copy(), t6["u_d_array"], initial_control(V, "ICT", c)))
    for K in (10, 100, 400):
        t = np.load(os.path.join(gold, f"traj_{K}_buoys.npz"))
        cases.append(Case(f"PL {K} buoys", t["x_0_array"][:, 0, :].copy(), t["u_d_array"], initial_control(V, "PL")))
    x0 = reference_grid()
    ocp = OCP(V, Parameters(), x0, np.zeros((x0.shape[0], NT, 2)), device=dev)
    ocp._primal(torch.from_numpy(golden_field()).to(dev), ocp.d_x, ocp.d_u, ocp.d_mask)
    ud = ocp._to_reference_layout(ocp.d_u)
    ocp.close()
    cases.append(Case("PL 10000 buoys", x0, ud, initial_control(V, "PL")))
    n_total = len(cases)
    cases = cases[rank::world]                 # independent cases: sharded over the ranks, no collective on the data path
    E = Ensemble(V, cases, device=dev)
    warm = max(args.warmup, 3)

    def timed(concurrent):
        for _ in range(warm):
            res = E.gradients(concurrent=concurrent)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = E.gradients(concurrent=concurrent)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.steps * 1e3, res

    ms_seq, r_seq = timed(False)
    if group is not None:
        dist.barrier(group=group)
    ms_con, r_con = timed(True)
    dJ = max(abs(a["J"] - b["J"]) / max(abs(a["J"]), 1e-300) for a, b in zip(r_seq, r_con))
    if group is not None:                      # whole-job time = the slowest rank
        t = torch.tensor([ms_seq, ms_con, dJ], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        ms_seq, ms_con, dJ = (float(v) for v in t.tolist())
    n = n_total
    if rank != 0:
        E.close()
        dist.destroy_process_group()
        return
    print(json.dumps({
        "metric": "ensemble_gd_iters_per_sec", "value": n / (ms_con * 1e-3), "unit": "GD iterations/s (aggregate over the cases)",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_con, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "reference fixtures (tests/golden) + synthetic 10000-buoy grid",
        "config": {"workload": "cfg4 ensemble: ICT cases 0-3 + PL 10/100/400/10000 buoys, square 32x32, one context/stream/thread per case",
                   "cases": [c.name for c in cases], "timing": "host wall clock around one gradient evaluation of every case"},
        "sequential_ms_per_step": ms_seq, "sequential_gd_iters_per_sec": n / (ms_seq * 1e-3),
        "concurrency_gain": ms_seq / ms_con, "max_rel_cost_difference_concurrent_vs_sequential": dJ,
        "J_rank0": [r["J"] for r in r_con], "newton_its_rank0": [r["newton_its"] for r in r_con]}), flush=True)
    E.close()
    if group is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true",
                    help="skip the 2^20-buoy kernel sweep, the per-function host calls and the 1e7-buoy sweep_strong leg")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "sweep", "ensemble"],
                    help="cfg3 (default, the headline GD iteration), the cfg5 synthetic drifter sweep, or the cfg4 ensemble of "
                         "independent cases run concurrently on one GPU")
    ap.add_argument("--sweep-buoys", type=int, default=10_000_000)
    ap.add_argument("--sweep-mesh", type=int, default=128,
                    help="x_resolution of the sweep mesh (cfg5: a refined mesh; 32 = the reference mesh and its stored field)")
    args = ap.parse_args()
    if args.workload == "ensemble" and args.impl == "ours":
        run_ensemble(args)
        return
    if args.workload == "sweep" and args.impl == "ours":
        run_sweep(args)
        return
    if args.impl == "reference":
        run_reference_arm(args, int(os.environ.get("RANK", "0")))
        return
    run_ours(args)


if __name__ == "__main__":
    main()
