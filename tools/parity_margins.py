"""Achieved parity errors of the CUDA path against the reference-source goldens, WITHOUT the conditioning floors of
tests/cases.py (plain per-derivative-block relative errors), plus the quirk-Q1 quantities (per-step local diffusion,
calibrated diffusion, rescaled final covariance of simulate_final_state).

    python tools/parity_margins.py [out.json]          # needs a GPU; writes profiles/r02_parity_margins.json

For every tests/golden/reference_*.npz trajectory file:
  mean_rel / cov_rel      max over steps of block_rel (no floors); *_excess = the test metric (floors on, < 1 passes)
  diff_local_rel          max over steps of |sigma^2_loc - golden| / golden   (solution_generator, step by step)
  diff_cal_rel            calibrated diffusion of solve()
  final_cov_rel           block_rel of L L^T of simulate_final_state (rescaled by the calibration) vs final_cov_sqrtm
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
from oracle import ek1_np  # noqa: E402


def product_case(g):
    prob, num, bcond = str(g["problem"]), int(g["num"]), str(g["bcond"])
    kw = dict(bcond=bcond) if prob in ("heat", "spruce") else {}
    return cases.make_case(prob, num=num, tmax=float(g["tmax"]), nu=int(g["nu"]), **kw)


def margins_for(path):
    g = np.load(path, allow_pickle=False)
    kind, nu, dt = str(g["kind"]), int(g["nu"]), float(g["dt"])
    n = nu + 1
    case = product_case(g)
    sol = cases.make_solver(kind, case).solve(case["pde"])
    with cases.perturbed_oracle():
        eps = ek1_np.solve(kind, case["opde"], dt, nu, case["gram_sqrtm"])
    mean, chol = sol.mean.cpu().numpy(), sol.cov_sqrtm.cpu().numpy()
    out = dict(kind=kind, D=int(chol.shape[-1]), steps=int(len(g["t"]) - 1))
    out["mean_rel"] = max(cases.block_rel(mean[k], g["mean"][k], n) for k in range(len(g["t"])))
    out["mean_rel_steps_only"] = max(cases.block_rel(mean[k], g["mean"][k], n) for k in range(1, len(g["t"])))
    out["cov_rel"] = max(cases.block_rel(cases.cov(chol[k]), cases.cov(g["cov_sqrtm"][k]), n) for k in range(len(g["t"])))
    out["mean_excess"] = max(cases.mean_excess(mean[k], g["mean"][k], spread=eps.mean[k]) for k in range(len(g["t"])))
    out["cov_excess"] = max(cases.cov_excess(chol[k], g["cov_sqrtm"][k], n) for k in range(len(g["t"])))
    # the reference's own reproducibility: eps-perturbed oracle vs golden (what any independent implementation can reach)
    out["mean_rel_eps_oracle"] = max(cases.block_rel(eps.mean[k], g["mean"][k], n) for k in range(len(g["t"])))
    out["cov_rel_eps_oracle"] = max(cases.block_rel(cases.cov(eps.cov_sqrtm[k]), cases.cov(g["cov_sqrtm"][k]), n)
                                    for k in range(len(g["t"])))
    cal = float(sol.diffusion_squared_calibrated)
    out["diff_cal_rel"] = abs(cal - float(g["diffusion_squared_calibrated"])) / abs(float(g["diffusion_squared_calibrated"]))
    out["diff_cal_rel_eps_oracle"] = abs(float(eps.diffusion_squared_calibrated) - float(g["diffusion_squared_calibrated"])) / abs(
        float(g["diffusion_squared_calibrated"]))
    states = [s for s, _ in cases.make_solver(kind, case).solution_generator(case["pde"])][1:]
    loc = np.array([float(s.diffusion_squared_local) for s in states])
    out["diff_local_rel"] = float(np.max(np.abs(loc - g["diffusion_squared_local"]) / np.abs(g["diffusion_squared_local"])))
    final, _ = cases.make_solver(kind, case).simulate_final_state(case["pde"])
    Lf = final.y.cov_sqrtm.cpu().numpy()
    out["final_cov_rel"] = cases.block_rel(cases.cov(Lf), cases.cov(g["final_cov_sqrtm"]), n)
    return {k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in out.items()}


def main():
    import __graft_entry__

    __graft_entry__.ensure_built()
    files = sorted(p for p in glob.glob(os.path.join(ROOT, "tests", "golden", "reference_*.npz"))
                   if "adaptive" not in p and "kalman" not in p)
    res = {}
    for fam in ("cta", "large"):
        os.environ["PNMOL_B200_PATH"] = fam
        for p in files:
            name = os.path.basename(p)[len("reference_"):-4]
            try:
                res[f"{fam}:{name}"] = margins_for(p)
            except Exception as exc:  # keep going: this is a report
                res[f"{fam}:{name}"] = {"error": repr(exc)[:300]}
            print(fam, name, json.dumps(res[f"{fam}:{name}"]), flush=True)
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    doc = {"what": "achieved parity errors vs reference-source goldens, conditioning floors OFF (block_rel); "
                   "*_excess are the test metrics (floors on, pass iff < 1); north-star tolerances: mean 1e-9, cov 1e-8",
           "git_head": head, "device": torch.cuda.get_device_name(0), "configs": res}
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_parity_margins.json")
    json.dump(doc, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
