"""Large-scene order (rp_batch_cfg.solve_order = RP_ORDER_COLOURED): the Gauss-Seidel sweeps walk a greedy colouring of
the constraint graph instead of the reference's array order. Not bit-comparable with the reference by construction; the
bar (BASELINE.json north_star) is physical: stacks settle, nothing tunnels, energy does not grow, and the result stays
within solver accuracy of the reference-order run. Bounds below are 3-10x the values measured on a B200 (printed)."""
import numpy as np
import pytest

import refdrv
import scenes

pytestmark = pytest.mark.gpu


def make(pkg, sc, coloured, n_worlds=1):
    b = pkg.Batch(pkg.Scene(sc), n_worlds=n_worlds, device=0, coloured=coloured)
    b.set_scene_forces(sc)
    return b


def energy(st, masses, g=10.0):
    """kinetic (linear) + potential energy of the non-fixed bodies; state records [n][21]"""
    v2 = (st[:, 7:10] ** 2).sum(axis=1)
    return float((0.5 * masses * v2 + masses * g * st[:, 1]).sum())


def test_coloured_wall_settles_with_few_colours(pkg):
    sc = scenes.brick_wall(rows=32, cols=32)
    masses = np.array([0.0 if b.fixed else b.mass for b in sc.bodies])
    ref, col = make(pkg, sc, False), make(pkg, sc, True)
    e0 = energy(col.state()[0], masses)
    y0 = col.state()[0][:, 1].copy()
    for f in range(90):
        ref.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        col.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    a, b = ref.state()[0], col.state()[0]
    cr, cc = ref.counters(), col.counters()
    depth_ref, depth_col = cr["levels"] / cr["frames"], cc["levels"] / cc["frames"]

    def summary(st):
        return dict(speed=float(np.sqrt((st[1:, 7:10] ** 2).sum(axis=1)).max()), sag=float((y0[1:] - st[1:, 1]).max()),
                    rise=float((st[1:, 1] - y0[1:]).max()), lowest=float(st[1:, 1].min()), energy=energy(st, masses))

    sa, sb = summary(a), summary(b)
    dy = float(np.abs(a[:, 1] - b[:, 1]).max())
    dxz = float(np.abs(a[:, [0, 2]] - b[:, [0, 2]]).max())
    print("wall after 90 frames: sweep depth %.0f levels (reference order) vs %.0f colours; between the two orders max |dy| %.3g, "
          "max lateral %.3g; initial energy %.6g\n  reference order: %r\n  coloured order:  %r" % (depth_ref, depth_col, dy, dxz, e0, sa, sb))
    assert not col.status().any() and not ref.status().any()
    assert depth_col <= 20 and depth_col * 4 <= depth_ref
    assert np.isfinite(b).all()
    # The wall drops 0.36 onto the floor and closes its 0.01 row gaps in both orders. Lateral drift of a 32-row dry wall is
    # chaotic (the two orders separate by decimetres within a second, as two reference builds with different rounding do,
    # SURVEY.md 8c), so the comparison is on what is determined: heights, the lowest brick, speeds, energy.
    assert abs(sb["sag"] - sa["sag"]) < 0.02 and abs(sb["rise"] - sa["rise"]) < 0.02 and dy < 0.05
    assert abs(sb["lowest"] - sa["lowest"]) < 0.01 and sb["lowest"] > -1.0 + 0.35 - 0.01   # on the floor's top face, not in it
    assert sb["speed"] < 2.0 * sa["speed"] + 0.1 and dxz < 1.0
    assert sb["energy"] <= e0 + 1e-6 * abs(e0) and abs(sb["energy"] - sa["energy"]) < 1e-3 * abs(sa["energy"])


def test_coloured_stack_tracks_reference(pkg, oracle_flavour):
    sc = scenes.stack()
    col = make(pkg, sc, True, n_worlds=3)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    worst_y = worst_xz = 0.0
    for f in range(120):
        col.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        got, want = col.state()[0], o.state()
        worst_y = max(worst_y, float(np.abs(got[:, 1] - want[:, 1]).max()))
        worst_xz = max(worst_xz, float(np.abs(got[:, [0, 2]] - want[:, [0, 2]]).max()))
    got = col.state()
    print("stack, coloured order vs oracle over 120 frames: worst |dy| = %.3g, worst lateral = %.3g" % (worst_y, worst_xz))
    assert (got == got[0]).all() and not col.status().any()
    # heights are determined (the stack lands and rests); the lateral creep of the reference's own stack (6.5 cm at the top
    # cube after 1 s, SURVEY.md 8c KAT) is order-sensitive
    assert worst_y < 0.05 and worst_xz < 0.5


def test_coloured_order_with_joints(pkg, oracle_flavour):
    """external constraints take part in the colouring (their colours are template constants)"""
    sc = scenes.hinge_joints()
    col = make(pkg, sc, True, n_worlds=2)
    col.broadcast(pkg.state15_to_21(sc.initial_state))
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(60):
        col.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    diff = float(np.abs(col.state()[0, :, :7] - o.state()[:, :7]).max())
    print("levers, coloured order vs oracle: |d pose| = %.3g" % diff)
    assert not col.status().any() and diff < 1e-2
