"""Phase breakdown of the step kernel (clock64 counters, see pnmol_b200_profile)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import bench
from pnmol_b200 import _lib, ensemble, kernels, white
from pnmol_b200.odetools import step
from pnmol_b200.pde import examples

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
T = int(sys.argv[2]) if len(sys.argv) > 2 else 48
if len(sys.argv) > 3 and sys.argv[3] == "sir17":   # the semilinear SIR ensemble of bench.py's side measurements
    pde = examples.sir_1d_discretized(num=17, tmax=T * bench.DT, diffusion_rate_S=0.035, diffusion_rate_I=0.035, diffusion_rate_R=0.035)
    solver = white.SemiLinearWhiteNoiseEK1(num_derivatives=2, steprule=step.Constant(bench.DT),
                                           spatial_kernel=kernels.duplicate(kernels.Matern52() + kernels.WhiteNoise(), 3))
    rng = np.random.default_rng(bench.SEED)
    y0 = np.tile(pde.y0, (M, 1)) * rng.uniform(0.9, 1.1, (M, 1))
    es = ensemble.EnsembleSolver(solver, pde, y0=y0, diff_scale=np.exp(rng.uniform(np.log(0.3), np.log(3.0), (M, 3))))
else:
    pde = examples.heat_1d_discretized(num=bench.NUM_POINTS, tmax=T * bench.DT, diffusion_rate=0.035)
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
names = ["mean+evaluate_ode", "build predict", "QR predict (rest)", "error estimate", "build update", "QR update (rest)", "solves+mean", "outputs",
         "qr: panel load", "qr: subpanel factor (1 warp)", "qr: subpanel apply", "qr: trailing apply", "qr: barrier after T",
         "update: right block (zero fill)", "qr: T factor", "update: error-estimate solve (warp 0)", "split: loads+pass1", "split: exchange barrier", "split: sum+T+pass2+stores", "-",
         "split: entry (row map, call)", "probe: mark-to-mark (overhead of a mark)", "probe: panel_rows", "(panel: rest = update+own -> in panel factor)"]
tot = float(out[:24].sum())
ms = e0.elapsed_time(e1)
print(f"members {M} steps {len(es.dts)}  kernel {ms:.1f} ms  -> {M*len(es.dts)/ms*1e3:.0f} member-steps/s")
for n, c in zip(names, out[:24]):
    print(f"  {n:20s} {100.0*float(c)/tot:6.2f} %   {float(c)/(M*len(es.dts)):12.0f} cycles / member-step")
