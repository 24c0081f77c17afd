"""CPU tests of the oracle itself (-m "not gpu").

Pinning chain: (1) the UNMODIFIED reference compiled into oracle/_ref/libref_oracle.so reproduces the known-answer
vector of SURVEY.md 8c; (2) the committed fixtures tests/golden/*.npz are outputs of that library (make_golden.py);
(3) the CPU restatement oracle/port is bit-identical to both. Where oracle/_ref is absent (GPU box without the prebuilt
file) the tests against it skip and the fixtures carry the pin.
"""
import os

import numpy as np
import pytest

import refdrv
import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "trajectories.npz"))
HULLS = np.load(os.path.join(HERE, "golden", "hulls.npz"))
WINDOWS = np.load(os.path.join(HERE, "golden", "windows.npz"))

need_ref = pytest.mark.skipif(not refdrv.available("strict"), reason="oracle/_ref/libref_oracle.so not present")

TRAJ = {"stack": {}, "brick_wall": {}, "cube_storm": {}, "seesaw": {}, "cube_and_ramp": {}, "coin": {}, "spring": {}, "hinge_joints": {},
        "arm": {}, "triple_pendula": {}, "rott_pendulum": {}, "mutual_orientation": {}, "negative_axes": {}, "mirror_cube": {}, "spheres": {}, "pile": dict(n_side=3), "tumble": {}, "w256": {}, "spot_storm": dict(n=2)}


def golden_frames(name):
    return sorted(int(k.split("/")[2]) for k in GOLD.files if k.startswith(name + "/state/"))


@pytest.fixture(scope="module", autouse=True)
def built(pkg):
    return pkg


@need_ref
def test_reference_known_answer_vector():
    """SURVEY.md 8c / BASELINE.md 4: unmodified stack.cpp scene, dt = 1/60, 20 substeps, strict flags."""
    w = refdrv.RefWorld("strict").load(scenes.stack())
    w.step()
    s = w.state()
    assert s[1, 0] == 1.7976086120357854e-07 and s[1, 1] == 1.3108716819988978e-06 and s[1, 2] == -7.513554353965761e-07
    assert s[2, 1] == 2.4985416666666804 and s[8, 1] == 17.498541666666867
    for _ in range(59):
        w.step()
    s = w.state()
    assert tuple(s[8, :3]) == (0.06548885409593784, 14.041351769773531, 0.029544658022551778)
    assert tuple(s[8, 3:7]) == (-0.0025164239475311307, 0.013251327303589749, -0.012339542698090008, 0.99983288884753707)
    assert tuple(s[4, :3]) == (-0.017016060596151589, 5.9999893644589051, -0.028480250088313414)


def test_port_known_answer_vector():
    w = refdrv.RefWorld("port").load(scenes.stack())
    for _ in range(60):
        w.step()
    s = w.state()
    assert tuple(s[8, :3]) == (0.06548885409593784, 14.041351769773531, 0.029544658022551778)
    assert tuple(s[1, :3]) == (-0.00035303731540455469, -5.8372691056387857e-06, 0.026016811197624966)


@pytest.mark.parametrize("name", sorted(TRAJ))
def test_port_matches_golden_trajectories(name):
    """Bit-exact: poses, velocities, sleep state at every recorded frame; static parameters; frame-1 contact log."""
    sc = scenes.BUILDERS[name](**TRAJ[name])
    w = refdrv.RefWorld("port").load(sc)
    assert np.array_equal(w.params(), GOLD[name + "/params"])
    w.log_enable(True)
    done = 0
    for f in golden_frames(name):
        while done < f:
            w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
            done += 1
        assert np.array_equal(w.state(), GOLD["%s/state/%d" % (name, f)]), (name, f)
        assert np.array_equal(w.prev_velocities(), GOLD["%s/prev_vel/%d" % (name, f)]), (name, f)
        if f == 1:
            calls, contacts = w.log_get()
            assert np.array_equal(calls, GOLD[name + "/calls"])
            assert np.array_equal(contacts, GOLD[name + "/contacts"])
            w.log_enable(False)
    assert w.lib.port_status() == 0


@need_ref
@pytest.mark.parametrize("name", ["stack", "tumble", "coin", "pile", "arm", "mirror_cube"])
def test_port_matches_reference_live(name):
    """Longer runs than the fixtures hold, port and reference stepped side by side, contact logs compared per frame."""
    sc = scenes.BUILDERS[name](**TRAJ[name])
    r = refdrv.RefWorld("strict").load(sc)
    p = refdrv.RefWorld("port").load(sc)
    r.log_enable(True)
    p.log_enable(True)
    for f in range(150):
        r.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        p.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        assert np.array_equal(r.state(), p.state()), (name, f)
        (rc, rk), (pc, pk) = r.log_get(), p.log_get()
        assert np.array_equal(rc, pc) and np.array_equal(rk, pk), (name, f)
        r.log_clear()
        p.log_clear()


@pytest.mark.parametrize("mesh", ["cube", "floor", "ico", "ramp", "cylinder", "lever", "seesaw_support"])
def test_hull_topology_matches_golden(mesh):
    """collider_convex_hull_create order (collider.cpp:194-364): vertices, faces, all three adjacency maps."""
    sc = scenes.Scene("h")
    sc.bodies.append(scenes.BodyDesc((0, 0, 0), scenes.IDENT, 1.0, False, [scenes.hull(mesh, (1.0, 1.0, 1.0))]))
    h = refdrv.RefWorld("port").load(sc).hull(0)
    for k, v in h.items():
        assert np.array_equal(v, HULLS["%s/%s" % (mesh, k)]), (mesh, k)


def test_cube_hull_is_the_documented_one():
    """cube.obj x (1.5,1,1) as the reference builds it."""
    sc = scenes.stack(1)
    h = refdrv.RefWorld("port").load(sc).hull(1)
    assert h["verts"].tolist() == [[-1.5, 1, -1], [1.5, 1, 1], [1.5, 1, -1], [-1.5, -1, 1], [1.5, -1, 1], [-1.5, 1, 1], [-1.5, -1, -1], [1.5, -1, -1]]
    assert h["face_idx"].reshape(6, 4).tolist() == [[5, 1, 2, 0], [5, 3, 4, 1], [0, 6, 3, 5], [4, 3, 6, 7], [1, 4, 7, 2], [2, 7, 6, 0]]
    assert h["v2f_idx"][h["v2f_ptr"][0]:h["v2f_ptr"][1]].tolist() == [0, 0, 2, 5, 5]
    assert h["v2n_idx"][h["v2n_ptr"][7]:h["v2n_ptr"][8]].tolist() == [3, 6, 2, 4, 0]


@need_ref
def test_pair_probes_match_reference():
    """Per-function known answers on random poses: GJK verdict + simplex, EPA normal/depth, manifold points."""
    rng = np.random.RandomState(3)
    shapes = [lambda: scenes.hull("cube", (1.0, 1.0, 1.0)), lambda: scenes.hull("ico", (1.0, 1.0, 1.0)),
              lambda: scenes.hull("cylinder", (1.0, 0.5, 1.0)), lambda: scenes.sphere(1.0), lambda: scenes.hull("ramp", (1.0, 1.0, 1.0))]
    hits = 0
    for trial in range(400):
        sc = scenes.Scene("probe")
        a, b = rng.randint(len(shapes)), rng.randint(len(shapes))
        if a == 3 and b == 3:
            b = 0
        for k, sh in enumerate((a, b)):
            q = scenes.quaternion_new(rng.rand(3), -180.0 + 360.0 * rng.rand())
            pos = tuple((rng.rand(3) - 0.5) * (1.6 if k else 0.0))
            sc.bodies.append(scenes.BodyDesc(pos, q, 1.0, False, [shapes[sh]()]))
        r = refdrv.RefWorld("strict").load(sc).probe_pair(0, 1)
        p = refdrv.RefWorld("port").load(sc).probe_pair(0, 1)
        assert r["hit"] == p["hit"] and r["epa_ok"] == p["epa_ok"]
        assert np.array_equal(r["simplex"], p["simplex"])
        assert np.array_equal(r["normal"], p["normal"]) and r["penetration"] == p["penetration"]
        assert np.array_equal(r["contacts"], p["contacts"])
        hits += int(r["hit"])
    assert hits > 100


@need_ref
def test_broadphase_pairs_match_reference():
    sc = scenes.pile(n_side=3, spacing=2.3)
    r = refdrv.RefWorld("strict").load(sc)
    p = refdrv.RefWorld("port").load(sc)
    for _ in range(30):
        r.step()
        p.step()
    assert np.array_equal(r.broad_pairs(), p.broad_pairs())
    assert len(r.broad_pairs()) > 27


def test_bounds_cull_is_sound_for_reference_gjk():
    """The CUDA path skips GJK for collider pairs whose world-space bounds are separated by > 1e-7 (k_cull). That is only
    exact if the reference-order GJK (with its quirks q2/q3) never reports such a pair as colliding: probed here on 1.5 M
    random box / octahedron pairs, a quarter axis-aligned, a seventh nearly touching."""
    import ctypes as C
    L = C.CDLL(refdrv.lib_path("port"))
    L.port_cull_soundness.restype = C.c_uint64
    L.port_cull_soundness.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    sep, hits = C.c_uint64(), C.c_uint64()
    assert L.port_cull_soundness(1500000, 2024, C.byref(sep), C.byref(hits)) == 0
    assert sep.value > 300000 and hits.value > 300000


@pytest.mark.skipif(not refdrv.available("strict"), reason="needs oracle/_ref (built where /root/reference is present)")
@pytest.mark.parametrize("name", ["stack", "tumble", "coin"])
def test_port_previous_velocities_match_reference(name):
    """Entity::previous_linear_velocity / previous_angular_velocity (entity.h:45-46) after pbd_simulate"""
    sc = scenes.BUILDERS[name]()
    a, b = refdrv.RefWorld("strict").load(sc), refdrv.RefWorld("port").load(sc)
    for f in range(40):
        a.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        b.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    assert np.array_equal(a.state(), b.state()) and np.array_equal(a.prev_velocities(), b.prev_velocities())
    assert np.abs(a.prev_velocities()).max() > 0


@pytest.mark.skipif(not refdrv.available("strict"), reason="needs oracle/_ref (built where /root/reference is present)")
def test_port_spot_storm_matches_reference():
    """spot_storm.cpp: compound bodies of 11 hulls (up to 529 vertices / 1000 faces): hull construction, the default inertia
    over several hulls, collider-pair expansion order and the narrowphase on large hulls, against the compiled reference"""
    sc = scenes.spot_storm(n=2)
    a, b = refdrv.RefWorld("strict").load(sc), refdrv.RefWorld("port").load(sc)
    assert np.array_equal(a.params(), b.params())
    for c in (0, 1, 10):
        ha, hb = a.hull(1, c), b.hull(1, c)
        assert all(np.array_equal(ha[k], hb[k]) for k in ha)
    for f in range(24):
        a.step()
        b.step()
    assert np.array_equal(a.state(), b.state())
    assert a.state()[1:, 1].min() < 3.0  # they have landed: contacts were solved


def window_digest(calls, contacts):
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(calls, dtype=np.uint32).tobytes())
    h.update(np.ascontiguousarray(contacts, dtype=np.float64).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


@pytest.mark.parametrize("name,kw,frames", [("w256", {}, 60), ("brick_wall_32x32", dict(rows=32, cols=32), 30)])
def test_port_matches_timed_windows(name, kw, frames):
    """The windows bench.py TIMES (W256 frames 0..59, brick wall 32x32 frames 0..29): per-frame narrowphase call and contact
    counts, states at three frames and the digest of the last frame's whole contact log, against the compiled reference's
    (tests/golden/windows.npz)."""
    sc = scenes.BUILDERS[name.split("_32")[0]](**kw)
    w = refdrv.RefWorld("port").load(sc)
    w.log_enable(True)
    want = WINDOWS[name + "/counts"]
    for f in range(frames):
        w.log_clear()
        w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        calls, contacts = w.log_get()
        assert (len(calls), len(contacts)) == tuple(want[f]), (name, f)
        key = "%s/state/%d" % (name, f + 1)
        if key in WINDOWS.files:
            assert np.array_equal(w.state(), WINDOWS[key]), key
    assert np.array_equal(window_digest(calls, contacts), WINDOWS[name + "/last_log_digest"])
    sub = refdrv.split_substeps(calls)
    assert np.array_equal(np.asarray(sub[-1], dtype=np.uint32), WINDOWS[name + "/last_substep_calls"])
    assert np.array_equal(contacts[int(sub[-1][0][3]):], WINDOWS[name + "/last_substep_contacts"])
