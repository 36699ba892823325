"""Short ensemble run for ncu captures: python tools/short_run.py [members] [steps]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")]
import torch
import bench
from pnmol_b200 import ensemble, kernels, white
from pnmol_b200.odetools import step
from pnmol_b200.pde import examples

M = int(sys.argv[1]) if len(sys.argv) > 1 else 592
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pde = examples.heat_1d_discretized(num=int(sys.argv[3]) if len(sys.argv) > 3 else bench.NUM_POINTS, tmax=T * bench.DT, diffusion_rate=0.035)
solver = white.LinearWhiteNoiseEK1(num_derivatives=2, steprule=step.Constant(bench.DT),
                                   spatial_kernel=kernels.SquareExponential() + kernels.WhiteNoise())
y0, diff, prior = bench.member_parameters(M, pde.mesh_spatial.points[:, 0], bench.SEED)
es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=diff, prior_scale=prior)
mean, chol, _ = es.initialize()
for _ in range(2):
    out = es.engine.run(pde.t0, es.dts, mean, chol)
torch.cuda.synchronize()
print("ok", int(out["status"].max()))
