"""Pace keeping of the persistent step kernel (csrc/ek1_kernels.cuh: k_run, pace_wait), as a host-side model.

After every step a CTA adds one to a counter and waits until the counter has reached its own running target, which it
advances by the number of CTAs that have a member in the current round: min(grid, batch - round * grid).  The model
below runs that protocol under adversarial schedules (any CTA that can make progress may be picked next) and checks that
every CTA finishes, i.e. that the targets are consistent for member counts that are not a multiple of the grid, and that
no CTA is ever more than one step ahead of another CTA of its round."""
import random

import pytest


def run_protocol(batch, grid, nsteps, seed):
    rng = random.Random(seed)
    counter = 0
    # per-CTA state: member index b, step s, running target, waiting flag
    cta = [dict(b=c, s=0, target=0, waiting=False, done=c >= batch) for c in range(grid)]
    steps_done = [[0] * nsteps for _ in range((batch + grid - 1) // grid)]
    guard = 0
    while not all(c["done"] for c in cta):
        runnable = [c for c in cta if not c["done"] and (not c["waiting"] or counter >= c["target"])]
        assert runnable, "deadlock: every unfinished CTA waits for a count that cannot be reached"
        c = rng.choice(runnable)
        if c["waiting"]:
            c["waiting"] = False
            c["s"] += 1
            if c["s"] == nsteps:
                c["s"] = 0
                c["b"] += grid
                if c["b"] >= batch:
                    c["done"] = True
            continue
        # one EK1 step of member b, then the pace point
        rnd = c["b"] // grid
        active = min(grid, batch - rnd * grid)
        steps_done[rnd][c["s"]] += 1
        # nobody of this round is more than one step ahead
        if c["s"] + 1 < nsteps:
            assert steps_done[rnd][c["s"] + 1] == 0 or steps_done[rnd][c["s"]] <= active
        if active > 1:
            c["target"] += active
            counter += 1
            c["waiting"] = True
        else:   # a round with a single member does not wait (k_run: pace_active > 1)
            c["s"] += 1
            if c["s"] == nsteps:
                c["s"] = 0
                c["b"] += grid
                if c["b"] >= batch:
                    c["done"] = True
        guard += 1
        assert guard < 10_000_000
    for rnd, row in enumerate(steps_done):
        assert row == [min(grid, batch - rnd * grid)] * nsteps
    return counter


@pytest.mark.parametrize("batch,grid,nsteps", [(7, 3, 4), (4, 4, 3), (5, 2, 2), (6, 6, 5), (13, 4, 3), (9, 8, 2), (1, 1, 3),
                                               (10, 3, 1), (296 * 2 + 5, 296, 2)])
def test_pace_targets_are_consistent(batch, grid, nsteps):
    g = min(grid, batch)
    for seed in range(3):
        total = run_protocol(batch, g, nsteps, seed)
        expect = sum(min(g, batch - r * g) * nsteps for r in range((batch + g - 1) // g) if min(g, batch - r * g) > 1)
        assert total == expect
