"""Dataflow Gauss-Seidel sweeps (pos_flow / vel_flow, rp_kernels.cuh): per-world counters of finished units instead of
grid-wide barriers between dependency levels. The per-world order of units is the level order either way, so the results
must be bit-identical to the barrier form (RP_FLOW=0) -- contact scenes, several positional iterations, joint scenes and one
large coloured scene with the dataflow form forced on (RP_FLOW=2) -- and to the oracle."""
import os

import numpy as np
import pytest

import refdrv

pytestmark = pytest.mark.gpu


def run(pkg, name, params=(), perturb=False, worlds=64, frames=10, flow=None, iters=None, **kw):
    old = os.environ.pop("RP_FLOW", None)
    if flow is not None:
        os.environ["RP_FLOW"] = str(flow)
    try:
        scene, desc = pkg.example(name, params, perturb=perturb)
        b = pkg.Batch(scene, n_worlds=worlds, device=0, **kw)
    finally:
        os.environ.pop("RP_FLOW", None)
        if old is not None:
            os.environ["RP_FLOW"] = old
    b.set_scene_forces(desc)
    for _ in range(frames):
        b.step(1.0 / 60.0, desc.substeps, desc.iters if iters is None else iters, desc.collisions)
    st = b.state()
    bits = int(np.bitwise_or.reduce(b.status()))
    b.close()
    return st, bits, desc


@pytest.mark.parametrize("name,params,perturb,frames,iters", [
    ("stack", (), False, 45, None),        # contacts from frame ~30
    ("stack", (), False, 40, 3),           # several positional iterations in one item sequence
    ("w256", (2, 2, 4), False, 50, None),
    ("cube_storm", (), False, 40, None),
    ("seesaw", (), False, 40, None),       # joints + contacts (dataflow forced on)
    ("hinge_joints", (), True, 30, None),
    ("triple_pendula", (), True, 15, None),
])
def test_dataflow_sweeps_equal_barrier_sweeps(pkg, name, params, perturb, frames, iters):
    flow, fbits, _ = run(pkg, name, params, perturb, worlds=70, frames=frames, flow=2, iters=iters)
    bar, bbits, _ = run(pkg, name, params, perturb, worlds=70, frames=frames, flow=0, iters=iters)
    assert fbits == 0 and bbits == 0
    assert np.array_equal(flow.view(np.uint64), bar.view(np.uint64))
    assert np.array_equal(flow[0], flow[-1])


def test_sweep_form_is_a_batch_setting(pkg):
    """rp_batch_cfg.sweep_form (1 = barriers, 2 = dataflow) without the environment: same bits"""
    scene, desc = pkg.example("stack")
    out = []
    for form in (1, 2):
        b = pkg.Batch(scene, n_worlds=66, device=0, sweep_form=form)
        b.set_scene_forces(desc)
        for _ in range(40):
            b.step()
        out.append(b.state())
        assert not b.status().any()
        b.close()
    assert np.array_equal(out[0].view(np.uint64), out[1].view(np.uint64))


def test_dataflow_default_matches_the_oracle(pkg, oracle_flavour):
    """70 worlds of the stack scene take the dataflow form by default; 60 frames against the oracle stepped alongside"""
    st, bits, desc = run(pkg, "stack", worlds=70, frames=60)
    o = refdrv.RefWorld(oracle_flavour).load(desc)
    for _ in range(60):
        o.step(substeps=desc.substeps, iters=desc.iters, collisions=desc.collisions)
    assert bits == 0
    assert np.array_equal(st[69][:, :15], o.state())


def test_dataflow_on_one_coloured_scene(pkg):
    """a single world (one counter for every unit) in the coloured order: same colours, same order, same bits"""
    flow, fbits, _ = run(pkg, "brick_wall", (8, 8), worlds=1, frames=40, flow=2, coloured=True)
    bar, bbits, _ = run(pkg, "brick_wall", (8, 8), worlds=1, frames=40, flow=0, coloured=True)
    assert fbits == 0 and bbits == 0
    assert np.array_equal(flow.view(np.uint64), bar.view(np.uint64))
