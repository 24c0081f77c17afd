"""Single-precision build (librawphys_b200_f32.so: the same sources with real = float), an EXPERIMENT, not a product mode. The
reference's trajectories are knife-edge sensitive to rounding (SURVEY.md TL;DR 3), so it is not compared bit for bit; what holds
-- and is checked here -- is the first two seconds: finite states, no flags, free fall within centimetres of the FP64
reference, bodies landing at their resting heights, nothing tunnels, joints hold. What does NOT hold is pinned too, so that
nobody mistakes the build for more than it is: a loaded stack left standing for 4 s comes apart (the reference's GJK / EPA /
clipping on exactly axis-aligned boxes is not robust in float)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    lib = os.path.join(ROOT, "raw-physics_b200", "librawphys_b200_f32.so")
    assert os.path.exists(lib)
    out = str(tmp_path_factory.mktemp("f32") / "f32.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "f32_worker.py"), out], capture_output=True, text=True,
                       env=dict(os.environ, RAWPHYS_B200_LIB=lib), timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(out)


def test_f32_states_are_finite_and_flag_free(runs):
    for k in runs.files:
        if "/state/" in k:
            assert np.isfinite(runs[k]).all(), k
        if k.endswith("/status"):
            # In float, perfectly aligned boxes give EPA exactly degenerate faces far more often than in double (coordinates round
            # to the SAME values): such a pair is flagged (EPA_DEGENERATE = 2, EPA_NO_CONVERGENCE = 4, where the reference would
            # assert) and treated as contact-free for that substep. Tolerated here, the physical criteria below are what counts;
            # any other flag (capacity, NaN, solver) is a failure.
            assert not (runs[k] & ~6).any(), (k, runs[k])
        if "/same/" in k:
            assert runs[k][0], k  # identical worlds stay identical (the arithmetic is deterministic)


def test_f32_free_fall_tracks_the_reference(runs):
    """the cubes still in the air after 10 frames (bodies 2..8; body 1 starts on the floor) against the FP64 reference. The bound
    is loose on purpose and says why single precision is hard HERE: XPBD derives velocities from position differences over
    h = 1/1200 s, and a float position of magnitude 17 moves in steps of 1.9e-6, i.e. the derived velocity in steps of 2.3e-3 m/s
    -- a random walk that reaches the centimetre within 200 substeps (measured: 1.4e-2)."""
    got, want = runs["stack/state/10"][2:, :7], GOLD["stack/state/10"][2:, :7]
    assert np.abs(got - want).max() <= 3e-2


def test_f32_stack_lands(runs):
    """stack.cpp: 8 cubes of height 2 dropped onto a floor whose top face is at y = -1. After one second every cube has landed
    on the one below (cube k at y = 2 k within the solver's slack) and is nearly at rest, as in the FP64 reference"""
    st = runs["stack/state/60"]
    y = np.sort(st[1:, 1])
    assert np.abs(y - 2.0 * np.arange(8)).max() < 0.1, y
    assert np.abs(st[1:, [0, 2]]).max() < 0.25
    assert np.sqrt((st[1:, 7:10] ** 2).sum(1)).max() < 1.0
    assert np.abs(y - np.sort(GOLD["stack/state/60"][1:, 1])).max() < 0.1


@pytest.mark.xfail(strict=True, reason="single precision: the loaded stack comes apart between 1 s and 4 s (DESIGN.md 7); pinned so that a fix is noticed")
def test_f32_stack_stays_standing(runs):
    st = runs["stack/state/240"]
    assert np.abs(np.sort(st[1:, 1]) - 2.0 * np.arange(8)).max() < 0.1


def test_f32_nothing_tunnels(runs):
    for k in ("w256/state/60", "w256/state/120", "wall/state/90", "coin/state/60", "spheres/state/120"):
        st = runs[k]
        moving = ~runs[k.split("/")[0] + "/fixed"]
        assert st[moving, 1].min() > -1.0, (k, st[moving, 1].min())  # body centres stay above the floor's top face (y = -1)
        assert np.sqrt((st[:, 7:10] ** 2).sum(1)).max() < 30.0, k


def test_f32_joints_hold(runs):
    """the levers spin about their hinges: the hinge anchors stay put (distance of each lever's centre from where it started)"""
    st60 = runs["levers/state/60"]
    assert np.abs(st60[:, :3]).max() < 50.0 and np.abs(np.linalg.norm(st60[:, 3:7], axis=1) - 1.0).max() < 1e-5
