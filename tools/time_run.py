"""Plain device timing of the C5 ensemble time loop: python tools/time_run.py [members] [steps] [repeats]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")]
import torch
import bench
from pnmol_b200 import ensemble, kernels, white
from pnmol_b200.odetools import step
from pnmol_b200.pde import examples

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
T = int(sys.argv[2]) if len(sys.argv) > 2 else 48
R = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pde = examples.heat_1d_discretized(num=int(os.environ.get("PNMOL_NUM", bench.NUM_POINTS)), tmax=T * bench.DT, diffusion_rate=0.035)
solver = white.LinearWhiteNoiseEK1(num_derivatives=2, steprule=step.Constant(bench.DT),
                                   spatial_kernel=kernels.SquareExponential() + kernels.WhiteNoise())
y0, diff, prior = bench.member_parameters(M, pde.mesh_spatial.points[:, 0], bench.SEED)
es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=diff, prior_scale=prior)
mean0, chol0, _ = es.initialize()
best = 1e30
for _ in range(R + 1):
    mean, chol = mean0.clone(), chol0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = es.engine.run(pde.t0, es.dts, mean, chol, flags=int(os.environ.get('PNMOL_RUN_FLAGS', '0'))); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"{os.environ.get('PNMOL_B200_LIB', 'default')}: members {M} steps {T} best {best:.1f} ms -> {M * T / best * 1e3:.0f} member-steps/s status {int(out['status'].max())}")
