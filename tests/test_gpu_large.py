"""GPU tests of the large-scene prologue (raw-physics_b200/csrc/rp_large.cuh): uniform-grid broadphase with ordered
compaction, union-find islands, parallel graph colouring -- forced on (rp_batch_cfg.large_scene = 2) for scenes small enough
for the oracle, so that every piece is compared with the reference (broad.cpp:6-29, :70-116; pbd.cpp:476-533)."""
import numpy as np
import pytest

import refdrv
import scenes

pytestmark = pytest.mark.gpu


def make(pkg, sc, n_worlds=1, **kw):
    b = pkg.Batch(pkg.Scene(sc), n_worlds=n_worlds, device=0, **kw)
    b.set_scene_forces(sc)
    if sc.initial_state is not None:
        b.broadcast(pkg.state15_to_21(sc.initial_state))
    return b


def test_grid_broadphase_pairs_equal_the_reference_on_a_2k_pile(pkg, oracle_flavour):
    """2197 bodies (13^3 ico / cylinder / sphere lattice) + floor, spacing 2.2 so that neighbours pair up from the start and
    more as the pile settles: the grid broadphase's pair list equals broad_get_collision_pairs' pair for pair, in order, at
    several frames; then the trajectory of the piled-up state is stepped on both sides, bit for bit."""
    sc = scenes.pile(n_side=13, spacing=2.2)
    assert len(sc.bodies) == 2198
    b = make(pkg, sc, n_worlds=2, large_scene=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    seen = []
    for f in range(1, 41):
        b.step()
        if f in (1, 15, 30, 40):
            st = b.state()
            assert np.array_equal(st[0], st[1])
            o.set_state(st[0, :, :15])
            want = o.broad_pairs()
            got = b.broad_pairs(0)
            assert np.array_equal(got, want), (f, len(got), len(want))
            seen.append(len(want))
    assert min(seen) > 2197 + 2000 and len(set(seen)) > 1  # the floor pairs with everything; thousands of neighbour pairs on top, changing as it moves
    # from the piled state on: both sides step the same frames (the oracle needs ~1 s per frame at this size)
    for f in range(3):
        b.step()
        o.step()
        assert np.array_equal(b.state()[0, :, :15], o.state()), f
    assert not b.status().any()


@pytest.mark.parametrize("name,frames", [("stack", 600), ("cube_storm", 120), ("hinge_joints", 60)])
def test_union_find_islands_and_sleeping(pkg, oracle_flavour, name, frames):
    """the large-scene islands (union-find over pairs and joints) + sleep bookkeeping against the reference, incl. falling
    asleep (stack: everything inactive after 600 frames) -- same bits as the per-world label propagation"""
    sc = scenes.BUILDERS[name]()
    b = make(pkg, sc, n_worlds=3, large_scene=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(frames):
        b.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if f % 50 == 49 or f == frames - 1:
            got, want = b.state()[0, :, :15], o.state()
            if sc.constraints:
                assert np.abs(got[:, :7] - want[:, :7]).max() <= 1e-9 and np.array_equal(got[:, 13], want[:, 13])
            else:
                assert np.array_equal(got, want), (name, f)
    if name == "stack":
        assert not b.state()[0, 1:, 13].any()
    assert not b.status().any()


@pytest.mark.parametrize("name,kw", [("brick_wall", dict(rows=12, cols=12)), ("pile", dict(n_side=6, spacing=2.2)), ("mutual_orientation", {})])
def test_parallel_colouring_is_a_valid_colouring(pkg, name, kw):
    """Jones-Plassmann colouring of the constraint graph (coloured order, large scene): no two units of one colour share a
    non-fixed body, every unit that is not skipped has a colour, colours stay few; and it is reproducible."""
    sc = scenes.BUILDERS[name](**kw)
    b = make(pkg, sc, coloured=True, large_scene=2)
    for _ in range(45):
        b.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    pairs, colours = b.pair_levels(0)
    pairs2, colours2 = b.pair_levels(0)
    assert np.array_equal(pairs, pairs2) and np.array_equal(colours, colours2)
    fixed = np.array([body.fixed for body in sc.bodies])
    assert len(pairs) > 0 and (colours[~(fixed[pairs[:, 0]] & fixed[pairs[:, 1]])] > 0).all()
    taken = set()
    for (a, c2), col in zip(pairs, colours):
        for body in (a, c2):
            if col > 0 and not fixed[body]:
                assert (int(body), int(col)) not in taken, (body, col)
                taken.add((int(body), int(col)))
    # joints hold their colours too: no contact unit may reuse one on the same body (checked through the dynamics below)
    assert colours.max() <= 40
    assert not b.status().any() and np.isfinite(b.state()).all()


def test_coloured_large_scene_matches_the_coloured_small_path_physically(pkg):
    """same acceptance as tests/test_gpu_coloured.py: the wall stays a wall, nothing tunnels, energy does not grow"""
    sc = scenes.brick_wall(rows=16, cols=16)
    a = make(pkg, sc, coloured=True, large_scene=1)
    b = make(pkg, sc, coloured=True, large_scene=2)
    for _ in range(90):
        a.step()
        b.step()
    sa, sb = a.state()[0], b.state()[0]
    assert np.abs(sa[1:, 1] - sb[1:, 1]).max() < 5e-3  # heights within 5 mm
    assert sb[1:, 1].min() > -1.0 + 0.3  # nothing below the floor's top face
    assert not a.status().any() and not b.status().any()
