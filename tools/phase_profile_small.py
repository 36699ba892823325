"""Phase breakdown of the small-state kernel (clock64 counters of warp 0 of every CTA): python tools/phase_profile_small.py [num] [members]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from pnmol_b200 import _lib, ensemble, kernels, white
from pnmol_b200.odetools import step
from pnmol_b200.pde import examples

num = int(sys.argv[1]) if len(sys.argv) > 1 else 6
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
pde = examples.heat_1d_discretized(num=num, tmax=bench.TMAX, diffusion_rate=0.035)
solver = white.LinearWhiteNoiseEK1(num_derivatives=2, steprule=step.Constant(bench.DT),
                                   spatial_kernel=kernels.SquareExponential() + kernels.WhiteNoise())
y0, diff, prior = bench.member_parameters(M, pde.mesh_spatial.points[:, 0], bench.SEED)
es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=diff, prior_scale=prior)
lib = _lib.load()
mean, chol, _ = es.initialize()
es.engine.run(pde.t0, es.dts, mean.clone(), chol.clone())
torch.cuda.synchronize()
_lib.check(lib.pnmol_b200_profile(es.engine.h, 1, None))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); es.engine.run(pde.t0, es.dts, mean, chol); e1.record(); torch.cuda.synchronize()
out = np.zeros(24, np.uint64)
_lib.check(lib.pnmol_b200_profile(es.engine.h, 0, _lib.ptr(out)))
names = ["mean+evaluate_ode", "build predict", "QR predict", "error estimate", "build update", "QR update", "solves+mean", "outputs"]
tot = float(out[:8].sum())
ms = e0.elapsed_time(e1)
print(f"path {es.engine.path} members {M} steps {len(es.dts)}  kernel {ms:.1f} ms -> {M*len(es.dts)/ms*1e3:.0f} member-steps/s")
for n, c in zip(names, out[:8]):
    print(f"  {n:20s} {100.0*float(c)/tot:6.2f} %")
