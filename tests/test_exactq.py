"""The reference's compile-time switch USE_QUATERNIONS_LINEARIZED_FORMULAS (pbd.cpp:16, :560-576; pbd_base_constraints.cpp:4,
:83-121, :187-225) turned OFF: orientation updates by axis-angle quaternions (libm sin / cos) instead of the linearised formula.
The arithmetic core carries both forms (-DRP_EXACT_QUATERNIONS); oracle/_ref/libref_oracle_exactq.so is the reference compiled
without the #define, tests/golden/exactq.npz its output (tests/golden/make_golden_exactq.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import refdrv
import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "exactq.npz"))
LIN = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
SCENES = sorted({k.split("/")[0] for k in GOLD.keys()})


def frames_of(name):
    return sorted(int(k.split("/")[2]) for k in GOLD.keys() if k.startswith(name + "/state/"))


@pytest.mark.parametrize("name", SCENES)
def test_restatement_without_linearised_formulas_is_bit_exact(pkg, name):
    """same libm on both sides (glibc): the CPU restatement reproduces the fixture bit for bit"""
    assert refdrv.available("port_exactq")
    sc = scenes.BUILDERS[name]()
    w = refdrv.RefWorld("port_exactq").load(sc)
    frames = frames_of(name)
    for f in range(1, max(frames) + 1):
        w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if f in frames:
            assert np.array_equal(w.state(), GOLD["%s/state/%d" % (name, f)]), (name, f)


def test_the_switch_changes_the_trajectory():
    assert not np.array_equal(GOLD["stack/state/60"], LIN["stack/state/60"][:, :15])
    assert np.abs(GOLD["stack/state/10"] - LIN["stack/state/10"][:, :15]).max() < 1e-3  # (free fall: the two forms barely differ)


def test_live_reference_matches_its_fixture():
    if not refdrv.available("strict_exactq"):
        pytest.skip("oracle/_ref/libref_oracle_exactq.so absent")
    sc = scenes.BUILDERS["stack"]()
    w = refdrv.RefWorld("strict_exactq").load(sc)
    for f in range(1, 41):
        w.step()
    assert np.array_equal(w.state(), GOLD["stack/state/40"])


@pytest.mark.gpu
def test_gpu_exact_quaternion_build(pkg, tmp_path):
    """librawphys_b200_exactq.so against the fixture. CUDA's sin / cos and glibc's may differ in the last ulp and every contact
    solve calls them in this mode, so the comparison carries a tolerance: 1e-9 absolute on poses up to the first contacts, 1e-6
    at frame 60 (resting contact amplifies an ulp); the worlds of the batch agree with each other bit for bit."""
    lib = os.path.join(ROOT, "raw-physics_b200", "librawphys_b200_exactq.so")
    assert os.path.exists(lib)
    out = str(tmp_path / "exactq_gpu.npz")
    specs = ["%s:%s" % (n, ",".join(str(f) for f in frames_of(n))) for n in SCENES]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "exactq_worker.py"), out] + specs, capture_output=True, text=True,
                       env=dict(os.environ, RAWPHYS_B200_LIB=lib), timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.load(out)
    worst = {}
    for name in SCENES:
        assert not got["%s/status" % name].any(), name
        for f in frames_of(name):
            k = "%s/state/%d" % (name, f)
            worst[k] = float(np.abs(got[k][:, :7] - GOLD[k][:, :7]).max())
    print("exact-quaternion build, worst |pose diff| per record:", worst)
    for k, v in worst.items():
        assert v <= (1e-9 if int(k.split("/")[2]) <= 40 else 1e-6), (k, v)
    # and it is the other formula: not the linearised trajectory
    assert np.abs(got["stack/state/60"][:, :7] - LIN["stack/state/60"][:, :7]).max() > 1e-6
