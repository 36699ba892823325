#!/usr/bin/env python
"""Benchmark of the EK1 filter loop (BASELINE.json metric: EK1 filter steps/s = batch x steps,
device-timed, plus FP64 roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--members M]

Workload (config.workload = "c5_heat_ensemble"): M = 4096 members PER GPU of the C1-sized heat
problem (1-D heat equation, Dirichlet, N = 50 mesh points, IWP(nu=2): D = 150, m = 52), linear
white-noise EK1, dt = 2^-4, tmax = 3 => 48 time steps per solve; per member a random initial
condition, diffusivity and prior output scale.  One bench "step" = one pass
of the persistent time-loop kernel over all members (M x 48 EK1 steps per GPU).

Printed JSON (one line, rank 0): see the task contract.  `value` is device-timed with the states
resident in HBM; `e2e` goes through the public host-buffer API (pinned host y0 -> H2D ->
initialize -> time loop -> rescale -> D2H of means and factors).  --impl reference times the
NumPy/LAPACK oracle (the reference cannot be installed: no jax in the image) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

NUM_POINTS = 50
NU = 2
DT = 2.0 ** -4
TMAX = 3.0
SEED = 20261018


# ----------------------------------------------------------------------------- workload
def member_parameters(num_members, x, seed):
    """Per-member IC a exp(-(x-x0)^2/w^2) sin(pi x), diffusivity scale and prior output scale."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.05, 0.2, num_members)
    x0 = rng.uniform(0.3, 0.7, num_members)
    w = rng.uniform(0.5, 1.5, num_members)
    y0 = a[:, None] * np.exp(-((x[None, :] - x0[:, None]) ** 2) / w[:, None] ** 2) * np.sin(np.pi * x[None, :])
    # diffusivity ~ logU(0.01, 0.1) relative to the nominal 0.035 of the discretised operator
    diff = np.exp(rng.uniform(np.log(0.01), np.log(0.1), num_members)) / 0.035
    prior = np.exp(rng.uniform(np.log(0.1), np.log(10.0), num_members))
    return y0, diff, prior


def ncu_traffic_record(name="r02_ncu_k_run_c5"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the step kernel per member-step, from the committed `ncu --set full`
    capture record profiles/<name>.json (written by tools/ncu_summary.py: carries the git head and the hash of the CUDA
    sources of the captured build, the capture command and the member-steps per launch).  None when the record is
    missing or was taken from other kernel sources than the ones this library is built from (stale capture)."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", name + ".json")))
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import ncu_summary

        if rec.get("csrc_sha256") != ncu_summary.csrc_sha256():
            return None
        return rec
    except Exception:
        return None


def work_model(D, m, d):
    """Algorithmic bytes / flops per member-step."""
    b_alg = 8.0 * (2 * (D * D + D) + d + 2)
    f_alg = (10.0 / 3.0) * D ** 3 + 4.0 * m * D ** 2 + 3.0 * m ** 2 * D
    return b_alg, f_alg


# ----------------------------------------------------------------------------- CPU arm
def _cpu_member_steps(args):
    """Worker: one member of the workload on one core (single-threaded BLAS), `nsteps` EK1 steps."""
    idx, nsteps = args
    from threadpoolctl import threadpool_limits

    from oracle import ek1_np, setup_np

    with threadpool_limits(limits=1):
        prob = setup_np.heat_1d(num=NUM_POINTS, tmax=TMAX, diffusion_rate=0.035)
        kern = setup_np.Sum(setup_np.SE(), setup_np.White())
        Lk = np.linalg.cholesky(setup_np.gram(kern, prob.points))
        y0, diff, prior = member_parameters(idx + 1, prob.points[:, 0], SEED)
        member = setup_np.with_member(prob, diff_scale=diff[idx], y0=y0[idx])
        st = ek1_np.white_initialize(member, NU, prior[idx] * Lk)
        t0 = time.perf_counter()
        for _ in range(nsteps):
            st = ek1_np.white_step(member, st, DT, NU, prior[idx] * Lk)
        return time.perf_counter() - t0, bool(np.isfinite(st.mean).all())


def cpu_throughput(nsteps, cores):
    """Ensemble throughput of the oracle: one process per core, each one member x nsteps steps."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_member_steps, [(i, nsteps) for i in range(cores)])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    assert all(r[1] for r in res)
    return cores * nsteps / busy, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nsteps = 6
    for _ in range(min(args.warmup, 1)):
        cpu_throughput(2, cores)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_throughput(nsteps, cores)
        vals.append(v)
    value = float(np.mean(vals))
    sample = f"{cores} members (one per core, single-thread BLAS) x {nsteps} EK1 steps per bench step"
    line = {
        "impl": "reference", "metric": "ek1_filter_steps_per_sec", "value": value, "unit": "member-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (time.perf_counter() - t0) / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 0),
        "cpu_baseline": {"value": value, "unit": "member-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "member-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = NumPy/LAPACK restatement of the reference (oracle/): jax/jaxlib/tornadox are not "
                "installable in this image, so the reference's own JAX path cannot run",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, nsteps_inner):
    return {"workload": "c5_heat_ensemble", "problem": "heat_1d dirichlet N=50", "solver": "LinearWhiteNoiseEK1 nu=2",
            "D": 150, "m": 52, "members_per_gpu": args.members, "time_steps_per_solve": nsteps_inner or 48,
            "dt": DT, "tmax": TMAX, "l2_policy": "inputs_larger_than_l2 (state = members x 180 KB)",
            "parallelism": f"ensemble members sharded over {args.gpus} GPU(s), no data-path collective"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7 or not (t_begin <= ts <= t_end + 0.2):
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
def run_b200_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    cpu_base = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, wall = cpu_throughput(12, cores)  # before CUDA is initialised in this process
        cpu_base = {"value": v, "unit": "member-steps/s", "cores": cores, "kind": "port",
                    "sample": f"{cores} members (one per core, single-thread BLAS) x 12 EK1 steps of the same workload, "
                              f"{wall:.1f} s wall incl. process start"}

    import torch
    import torch.distributed as dist

    import __graft_entry__
    __graft_entry__.ensure_built()
    from pnmol_b200 import _lib, ensemble, kernels, white
    from pnmol_b200.odetools import step
    from pnmol_b200.pde import examples

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pde = examples.heat_1d_discretized(num=NUM_POINTS, tmax=TMAX, diffusion_rate=0.035)
    solver = white.LinearWhiteNoiseEK1(num_derivatives=NU, steprule=step.Constant(DT),
                                       spatial_kernel=kernels.SquareExponential() + kernels.WhiteNoise())
    M = args.members
    y0, diff, prior = member_parameters(M, pde.mesh_spatial.points[:, 0], SEED + rank)
    es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=diff, prior_scale=prior, device=dev)
    eng = es.engine
    T = len(es.dts)
    D, m, d = eng.D, eng.m, eng.d
    b_alg, f_alg = work_model(D, m, d)

    mean0, chol0, status0 = es.initialize()
    mean, chol = torch.empty_like(mean0), torch.empty_like(chol0)
    gathered = torch.empty((world * M, D), dtype=torch.float64, device=dev) if world > 1 else None

    def bench_step():
        mean.copy_(mean0)
        chol.copy_(chol0)
        out = eng.run(pde.t0, es.dts, mean, chol)
        if world > 1:  # results gathered over NVLink once per solve (means only; factors stay sharded)
            dist.all_gather_into_tensor(gathered, mean.reshape(M, D))
        return out

    for _ in range(max(args.warmup, 3)):
        out = bench_step()
    barrier()
    assert int(out["status"].max()) == 0 and int(status0.max()) == 0, "non-finite member"

    # kernel-only duration of the dominant kernel (k_run), CUDA events on the launching stream
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.launch_count()
    barrier()
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        mean.copy_(mean0)
        chol.copy_(chol0)
        kev[k][0].record()
        out = eng.run(pde.t0, es.dts, mean, chol)
        kev[k][1].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, mean.reshape(M, D))
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = _lib.launch_count() - launches0
    elapsed_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    if world > 1:
        tmax_ms = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax_ms, op=dist.ReduceOp.MAX)
        elapsed_ms, kernel_ms = float(tmax_ms[0]), float(tmax_ms[1])
    member_steps = world * M * T * args.steps
    value = member_steps / (elapsed_ms * 1e-3)

    # Strong scaling (BASELINE config 5: ONE ensemble of 4096 members sharded over the GPUs): the same time loop on this
    # rank's slice of args.members members, timed the same way (device events, max over ranks), incl. the final gather.
    strong = None
    if world > 1:
        from pnmol_b200.ensemble import member_slice

        sl = member_slice(M, world, rank)
        y0s, diffs, priors = member_parameters(M, pde.mesh_spatial.points[:, 0], SEED)
        es_s = ensemble.EnsembleSolver(solver, pde, y0=y0s[sl], diff_scale=diffs[sl], prior_scale=priors[sl], device=dev)
        m0s, c0s, _ = es_s.initialize()
        ms_, cs_ = torch.empty_like(m0s), torch.empty_like(c0s)
        Ms = sl.stop - sl.start
        gath_s = torch.empty((world * ((M + world - 1) // world), D), dtype=torch.float64, device=dev)
        pad = torch.zeros(((M + world - 1) // world, D), dtype=torch.float64, device=dev)

        def strong_step():
            ms_.copy_(m0s); cs_.copy_(c0s)
            o = es_s.engine.run(pde.t0, es_s.dts, ms_, cs_)
            pad[:Ms] = ms_.reshape(Ms, D)
            dist.all_gather_into_tensor(gath_s, pad)
            return o

        for _ in range(3):
            strong_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            strong_step()
        s1.record()
        barrier()
        tt = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "members_total": M, "members_per_gpu": Ms, "value": M * T * args.steps / (float(tt[0]) * 1e-3),
                  "unit": "member-steps/s", "ms_per_step": float(tt[0]) / args.steps,
                  "note": "one 4096-member ensemble sharded over the GPUs; members are indivisible units of 48 sequential "
                          "steps, so 4096 / N members per GPU on 296 resident CTAs run in ceil(4096 / N / 296) waves"}
        del es_s, m0s, c0s, ms_, cs_

    # FP64 peak: cuBLAS DGEMM through torch (same method as MEASURED_PEAKS.json uses for bf16)
    n_gemm = 4096
    a = torch.randn(n_gemm, n_gemm, dtype=torch.float64, device=dev)
    bmat = torch.randn(n_gemm, n_gemm, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ bmat
    best = 1e9
    for _ in range(5):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); a @ bmat; g1.record(); torch.cuda.synchronize()
        best = min(best, g0.elapsed_time(g1))
    fp64_peak = 2.0 * n_gemm ** 3 / (best * 1e-3) * 1e-12
    del a, bmat

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ach_tf = f_alg * M * T / (kernel_ms * 1e-3) * 1e-12
    ach_gbs = b_alg * M * T / (kernel_ms * 1e-3) * 1e-9
    rec = ncu_traffic_record()
    roofline = {"bound": "tensor", "pipe": "fp64 (mma.sync DMMA + DFMA; tcgen05 has no FP64 MMA)", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                "traffic": rec["dram_bytes_per_member_step"] * M * T if rec else None,
                "traffic_source": (f"ncu dram__bytes_read.sum + dram__bytes_write.sum per member-step x member-steps per launch; "
                                   f"capture of build {rec['git_head']} ({rec['command']}, profiles/r02_ncu_k_run_c5_summary.csv)"
                                   if rec else "no capture record of this build (profiles/r02_ncu_k_run_c5.json missing or stale)"),
                "kernel": "pnmol::k_run", "kernel_ms_per_launch": kernel_ms,
                "algorithmic_flops_per_member_step": f_alg, "algorithmic_bytes_per_member_step": b_alg,
                "peak_source": "measured in this run: cuBLAS DGEMM 4096^3 via torch.matmul(float64), best of 5 "
                               "(tools/fp64_peak.cu measured DMMA 37.2 / DFMA 34.0 TFLOP/s on this pool)",
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"}}

    # end to end through the public host-buffer API
    y0_pin = torch.from_numpy(es.y0).pin_memory()
    mean_pin = torch.empty((M, eng.n, eng.dd), dtype=torch.float64).pin_memory()
    chol_pin = torch.empty((M, D, D), dtype=torch.float64).pin_memory()
    es.y0 = y0_pin
    es.simulate_final_state_host(mean_host=mean_pin, chol_host=chol_pin)  # warm-up (allocates the host-path buffers)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = es.simulate_final_state_host(mean_host=mean_pin, chol_host=chol_pin)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
    assert int(res.status.max()) == 0
    e2e = {"value": world * M * T * e2e_steps / e2e_s, "unit": "member-steps/s", "h2d_bytes_per_step": int(M * d * 8),
           "d2h_bytes_per_step": int(M * (D + D * D + 1) * 8 + M * 4),
           "includes": "H2D of y0, initialize (2 QR updates), 48-step time loop, rescale, D2H of means + factors"}

    other = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        other = other_configs_timing(dev, fp64_peak)
        # the other BASELINE configs' roofline fractions next to the headline's (same denominators)
        roofline["other_configs"] = {k: {kk: v[kk] for kk in ("fp64_frac", "hbm_frac", "ms_per_step", "member_steps_per_sec", "path", "dram_over_algorithmic")
                                         if kk in v} for k, v in other.items() if "error" not in v}

    if rank == 0:
        line = {
            "metric": "ek1_filter_steps_per_sec", "value": value, "unit": "member-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, T), "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "strong_scaling": strong, "other_configs": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def other_configs_timing(dev, fp64_peak):
    """BASELINE configs C2-C4 (single large solves, multi-CTA kernels): device-timed ms per EK1 step, and the other
    C5 ensembles (SIR N=17, the N=6 probe): member-steps/s.  Not part of `value`; everything here stays below ~15 s."""
    import torch

    import cases

    configs = {"c2_sir_N100_white_semilinear": ("sir", "white_semilinear", "neumann", 100, 2.0 ** -3, "matern"),
               "c3_spruce_N200_latent_semilinear": ("spruce", "latent_semilinear", "dirichlet", 200, 2.0 ** -4, "se"),
               "c4_heat_N1024_white_linear": ("heat", "white_linear", "dirichlet", 1024, 2.0 ** -4, "se")}
    out = {}
    for name, (pname, kind, bcond, num, dt, prior) in configs.items():
        try:
            steps = 4 if num < 1000 else 2
            case = cases.make_case(pname, num=num, bcond=bcond, dt=dt, prior=prior, tmax=steps * dt)
            solver = cases.make_solver(kind, case)
            s0 = solver.initialize(case["pde"])
            eng = solver._engine
            D, m = eng.D, eng.m
            mean, chol = s0.y.mean.clone().reshape(1, eng.n, eng.dd), s0.y.cov_sqrtm.clone().reshape(1, D, D)
            dts = np.full(steps, dt)
            eng.run(case["pde"].t0, dts[:1], mean.clone(), chol.clone())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = eng.run(case["pde"].t0, dts, mean, chol)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            f_alg = work_model(D, m, eng.d)[1]
            b_alg = work_model(D, m, eng.d)[0]
            rec = ncu_traffic_record("r02_ncu_k_run_large_" + name[:2])
            out[name] = {"D": D, "m": m, "path": eng.path, "ms_per_step": ms, "steps_per_sec": 1e3 / ms,
                         "fp64_frac": f_alg / (ms * 1e-3) * 1e-12 / fp64_peak, "status": int(res["status"].max()),
                         "algorithmic_bytes_per_step": b_alg,
                         "dram_bytes_per_step": rec["dram_bytes_per_member_step"] if rec else None,
                         "dram_over_algorithmic": rec["dram_bytes_per_member_step"] / b_alg if rec else None}
            del solver, eng, mean, chol, s0
            torch.cuda.empty_cache()
        except Exception as exc:  # never lose the headline line to a side measurement
            out[name] = {"error": repr(exc)[:200]}
    # the other ensemble members (SIR N=17: D=153, m=57; the HBM-bound probe N=6: D=18, m=8)
    from pnmol_b200 import ensemble, kernels, white
    from pnmol_b200.odetools import step
    from pnmol_b200.pde import examples

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for name in ("c5_sir_N17_ensemble_4096", "c5_heat_N6_probe_ensemble_65536", "c5_heat_N12_ensemble_16384", "c5_heat_N24_ensemble_4096"):
        try:
            M = int(name.rsplit("_", 1)[1])
            rng = np.random.default_rng(SEED)
            if "sir" in name:
                pde = examples.sir_1d_discretized(num=17, tmax=TMAX, diffusion_rate_S=0.035, diffusion_rate_I=0.035,
                                                  diffusion_rate_R=0.035)
                solver = white.SemiLinearWhiteNoiseEK1(num_derivatives=NU, steprule=step.Constant(DT),
                                                       spatial_kernel=kernels.duplicate(kernels.Matern52() + kernels.WhiteNoise(), 3))
                es = ensemble.EnsembleSolver(solver, pde, y0=np.tile(pde.y0, (M, 1)) * rng.uniform(0.9, 1.1, (M, 1)),
                                             diff_scale=np.exp(rng.uniform(np.log(0.3), np.log(3.0), (M, 3))), device=dev)
            else:
                pde = examples.heat_1d_discretized(num=int(name.split("_N")[1].split("_")[0]), tmax=TMAX, diffusion_rate=0.035)
                solver = white.LinearWhiteNoiseEK1(num_derivatives=NU, steprule=step.Constant(DT),
                                                   spatial_kernel=kernels.SquareExponential() + kernels.WhiteNoise())
                y0, diff, prior = member_parameters(M, pde.mesh_spatial.points[:, 0], SEED)
                es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=diff, prior_scale=prior, device=dev)
            eng = es.engine
            mean0, chol0, _ = es.initialize()
            mean, chol = mean0.clone(), chol0.clone()
            eng.run(pde.t0, es.dts, mean, chol)
            mean.copy_(mean0); chol.copy_(chol0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = eng.run(pde.t0, es.dts, mean, chol)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            b_alg, f_alg = work_model(eng.D, eng.m, eng.d)
            rate = M * len(es.dts) / (ms * 1e-3)
            out[name] = {"D": eng.D, "m": eng.m, "path": eng.path, "member_steps_per_sec": rate,
                         "fp64_frac": rate * f_alg * 1e-12 / fp64_peak,
                         "hbm_frac": rate * b_alg * 1e-9 / peaks.get("hbm_gbs", 6650.0), "status": int(res["status"].max())}
            del es, eng, mean, chol, mean0, chol0
            torch.cuda.empty_cache()
        except Exception as exc:
            out[name] = {"error": repr(exc)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=4096, help="ensemble members per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C2-C4 side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
