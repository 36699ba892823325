"""Host-side logic of the multi-GPU ensemble: member partitioning and the final gather,
run with the gloo backend on CPU (world size 2) and a stub in place of the CUDA compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pnmol_b200 import ensemble


def test_member_slice_partitions_exactly():
    for B in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s = ensemble.member_slice(B, world, r)
                seen.extend(range(s.start, s.stop))
            assert seen == list(range(B))
            sizes = [ensemble.member_slice(B, world, r).stop - ensemble.member_slice(B, world, r).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _stub(solver, pde, *, y0, diff_scale, prior_scale, reaction_params):
    y0 = torch.as_tensor(y0)
    scale = torch.as_tensor(diff_scale).reshape(-1, 1) if diff_scale is not None else 1.0
    mean = (y0 * scale)[:, None, :].repeat(1, 3, 1)
    chol = torch.eye(3 * y0.shape[1], dtype=torch.float64)[None].repeat(y0.shape[0], 1, 1) * y0[:, :1, None]
    return ensemble.EnsembleResult(1.0, mean, chol, y0.sum(dim=1), torch.zeros(y0.shape[0], dtype=torch.int32), 4)


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y0 = np.arange(B * 4, dtype=np.float64).reshape(B, 4)
        ds = 1.0 + np.arange(B, dtype=np.float64)
        res = ensemble.simulate_final_state_distributed(None, None, y0=y0, diff_scale=ds, compute=_stub)
        full = _stub(None, None, y0=y0, diff_scale=ds, prior_scale=None, reaction_params=None)
        ok = (torch.equal(res.mean, full.mean) and torch.equal(res.cov_sqrtm, full.cov_sqrtm)
              and torch.equal(res.diffusion_squared_calibrated, full.diffusion_squared_calibrated))
        out.put((rank, bool(ok), tuple(res.mean.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 7])
def test_distributed_gather_gloo_world2(B):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)
    assert all(shape == (B, 3, 4) for _, _, shape in results)


# ------------------------------------------------------------------------- NCCL, 2 GPUs (runs only on a multi-GPU box)
def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys

        here = os.path.dirname(os.path.abspath(__file__))
        sys.path[:0] = [here]
        import cases

        case = cases.make_case("heat", num=9, tmax=0.25)
        pde = case["pde"]
        B = 7  # uneven split: 4 + 3 members
        x = pde.mesh_spatial.points[:, 0]
        y0 = np.stack([(0.05 + 0.02 * b) * np.exp(-((x - 0.5) ** 2)) * np.sin(np.pi * x) for b in range(B)])
        ds = 0.5 + 0.25 * np.arange(B)
        solver = cases.make_solver("white_linear", case)
        res = ensemble.simulate_final_state_distributed(solver, pde, y0=y0, diff_scale=ds)   # means + factors all-gathered
        full = ensemble.simulate_final_state(cases.make_solver("white_linear", case), pde, y0=y0, diff_scale=ds)
        ok = (res.mean.shape == full.mean.shape and torch.equal(res.mean, full.mean) and torch.equal(res.cov_sqrtm, full.cov_sqrtm)
              and torch.equal(res.diffusion_squared_calibrated, full.diffusion_squared_calibrated))
        out.put((rank, bool(ok), tuple(res.cov_sqrtm.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_distributed_gather_nccl_world2():
    """simulate_final_state_distributed over NCCL: every rank ends up with all members' means AND factors, bitwise equal
    to the single-GPU ensemble (one all-gather after the loop, no collective inside a step)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    assert all(shape == (7, 27, 27) for _, _, shape in results)
