"""Per-step timing of the multi-CTA path on BASELINE configs C2-C4 (CUDA events around attempt_step / run).
python tools/large_timing.py [c2|c3|c4 ...] [--steps K] [--check]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import cases

CONFIGS = {"c2": ("sir", "white_semilinear", "neumann", 100, 2.0 ** -3, "matern"),
           "c3": ("spruce", "latent_semilinear", "dirichlet", 200, 2.0 ** -4, "se"),
           "c4": ("heat", "white_linear", "dirichlet", 1024, 2.0 ** -4, "se"),
           "c1": ("heat", "white_linear", "dirichlet", 50, 2.0 ** -4, "se")}
args = [a for a in sys.argv[1:] if a in CONFIGS] or ["c2", "c3"]
steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 4
for name in args:
    pname, kind, bcond, num, dt, prior = CONFIGS[name]
    t0 = time.perf_counter()
    case = cases.make_case(pname, num=num, bcond=bcond, dt=dt, prior=prior, tmax=steps * dt)
    solver = cases.make_solver(kind, case)
    t_setup = time.perf_counter() - t0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); s0 = solver.initialize(case["pde"]); e[1].record()
    eng = solver._engine
    D, m = eng.D, eng.m
    mean, chol = s0.y.mean.clone().reshape(1, eng.n, eng.dd), s0.y.cov_sqrtm.clone().reshape(1, D, D)
    dts = np.full(steps, dt)
    eng.run(case["pde"].t0, dts[:1], mean.clone(), chol.clone())   # warm-up
    from pnmol_b200 import _lib
    prof = "--profile" in sys.argv
    if prof:
        _lib.check(eng.lib.pnmol_b200_profile(eng.h, 1, None))
    e[2].record(); out = eng.run(case["pde"].t0, dts, mean, chol); e[3].record()
    torch.cuda.synchronize()
    ms_init, ms_step = e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]) / steps
    flop = (10.0 / 3.0) * D ** 3 + 4.0 * m * D * D + 3.0 * m * m * D
    print(f"{name}: {pname} N={num} {kind} D={D} m={m} path={eng.path} host set-up {t_setup:.1f} s | initialize {ms_init:.2f} ms | "
          f"step {ms_step:.3f} ms = {1e3 / ms_step:.1f} steps/s | F_alg {flop / 1e9:.2f} Gflop -> {flop / ms_step / 1e9:.2f} TFLOP/s | status {int(out['status'].max())}")
    if prof:
        cyc = np.zeros(24, np.uint64)
        _lib.check(eng.lib.pnmol_b200_profile(eng.h, 0, _lib.ptr(cyc)))
        names = ["mean+evaluate_ode (CTA 0)", "build predict + barrier", "-", "error estimate", "build update", "-", "solves+mean+factor out", "end barrier",
                 "qr: panel factor (CTA 0)", "qr: barrier after panel", "qr: partial Y", "qr: barrier", "qr: update", "qr: barrier",
                 "panel: load", "panel: first norm", "panel: column loop", "panel: write-back + V", "panel: streamed apply", "panel: Gram + T"]
        tot = float(cyc.sum())
        for nme, c in zip(names, cyc[:20]):
            if c:
                print(f"    {nme:28s} {100.0 * float(c) / tot:6.2f} %  {float(c) / steps / 1.965e6:9.3f} ms/step")
