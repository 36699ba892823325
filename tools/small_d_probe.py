"""Throughput of the ensemble kernels on small members (HBM-bound probe at N=6, D=18; SIR N=17, D=153).
PNMOL_B200_PATH=cta|warp python tools/small_d_probe.py [heat6|sir17|heat12|heat24] [members]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from pnmol_b200 import ensemble, kernels, white
from pnmol_b200.odetools import step
from pnmol_b200.pde import examples

which = sys.argv[1] if len(sys.argv) > 1 else "heat6"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
if which.startswith("heat"):
    num = int(which[4:])
    pde = examples.heat_1d_discretized(num=num, tmax=bench.TMAX, diffusion_rate=0.035)
    solver = white.LinearWhiteNoiseEK1(num_derivatives=2, steprule=step.Constant(bench.DT),
                                       spatial_kernel=kernels.SquareExponential() + kernels.WhiteNoise())
    y0, diff, prior = bench.member_parameters(M, pde.mesh_spatial.points[:, 0], bench.SEED)
    es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=diff, prior_scale=prior)
else:
    pde = examples.sir_1d_discretized(num=17, tmax=bench.TMAX, diffusion_rate_S=0.035, diffusion_rate_I=0.035, diffusion_rate_R=0.035)
    solver = white.SemiLinearWhiteNoiseEK1(num_derivatives=2, steprule=step.Constant(bench.DT),
                                           spatial_kernel=kernels.duplicate(kernels.Matern52() + kernels.WhiteNoise(), 3))
    rng = np.random.default_rng(bench.SEED)
    y0 = np.tile(pde.y0, (M, 1)) * rng.uniform(0.9, 1.1, (M, 1))
    es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=np.exp(rng.uniform(np.log(0.3), np.log(3.0), (M, 3))))
eng = es.engine
mean0, chol0, st0 = es.initialize()
mean, chol = mean0.clone(), chol0.clone()
eng.run(pde.t0, es.dts, mean, chol)
best = 1e9
for _ in range(3):
    mean.copy_(mean0); chol.copy_(chol0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.run(pde.t0, es.dts, mean, chol); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
D, m, d = eng.D, eng.m, eng.d
b_alg, f_alg = bench.work_model(D, m, d)
T = len(es.dts)
rate = M * T / (best * 1e-3)
print(f"{which} path={eng.path} members {M} D={D} m={m}: {best:.1f} ms -> {rate:,.0f} member-steps/s | "
      f"{rate * b_alg / 1e9:.1f} GB/s algorithmic ({100 * rate * b_alg / 6454.6e9:.2f} % of HBM), "
      f"{rate * f_alg / 1e12:.3f} TFLOP/s ({100 * rate * f_alg / 35.4e12:.2f} % of FP64) | status {int(out['status'].max())}")
