"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same inputs.

Bar: BIT-EXACT. All arithmetic on the path is IEEE FP64 +,-,*,/,sqrt (plus a few float round trips the reference has),
compiled with --fmad=false, in the reference's operation order, so poses, velocities, sleep state, contact-pair sets,
manifold counts and contact points must be identical to the last bit -- for contact scenes. Joint-limit scenes call libm
asin/sin/cos (pbd.cpp:177, quaternion.cpp:7-10) where CUDA's and glibc's last ulp may differ: for those the stated
tolerance is 1e-9 absolute on positions/quaternions after the tested horizon (measured: see the test).
"""
import os

import numpy as np
import pytest

import refdrv
import scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
WINDOWS = np.load(os.path.join(ROOT, "tests", "golden", "windows.npz"))

CONTACT_SCENES = {"stack": ({}, 90), "brick_wall": ({}, 60), "cube_storm": ({}, 90), "seesaw": ({}, 120), "cube_and_ramp": ({}, 120),
                  "coin": ({}, 90), "mirror_cube": ({}, 120), "spheres": ({}, 150), "pile": (dict(n_side=3), 60), "tumble": ({}, 120),
                  "spring": ({}, 60),
                  # positional + mutual-orientation joints + contacts: no libm call on the path (pbd.cpp:156-173), so bit-exact
                  "mutual_orientation": ({}, 90)}
JOINT_SCENES = {"hinge_joints": ({}, 120), "arm": ({}, 120), "triple_pendula": ({}, 40),
                # a hinge limited to [0, 0] at 50 substeps x 50 iterations (rott_pendulum.cpp:154); NEGATIVE_* axis selectors (q4)
                "rott_pendulum": ({}, 30), "negative_axes": ({}, 60)}


def make(pkg, sc, n_worlds=1, **kw):
    b = pkg.Batch(pkg.Scene(sc), n_worlds=n_worlds, device=0, **kw)
    b.set_scene_forces(sc)
    if sc.initial_state is not None:
        b.broadcast(pkg.state15_to_21(sc.initial_state))
    return b


def step(b, sc):
    b.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)


@pytest.mark.parametrize("name", sorted(CONTACT_SCENES))
def test_trajectory_bit_exact(pkg, oracle_flavour, name):
    kw, frames = CONTACT_SCENES[name]
    sc = scenes.BUILDERS[name](**kw)
    b = make(pkg, sc, n_worlds=3)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(frames):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if f % 10 == 9 or f < 3 or f == frames - 1:
            got = b.state()
            want = o.state()
            assert np.array_equal(got[0, :, :15], want), (name, f, np.abs(got[0, :, :15] - want).max())
            assert np.array_equal(got[1], got[0]) and np.array_equal(got[2], got[0])
    assert not b.status().any()


@pytest.mark.parametrize("name", sorted(JOINT_SCENES))
def test_joint_scene_within_tolerance(pkg, oracle_flavour, name):
    """libm-dependent scenes: report bitwise equality when it holds, require 1e-9."""
    kw, frames = JOINT_SCENES[name]
    sc = scenes.BUILDERS[name](**kw)
    b = make(pkg, sc, n_worlds=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    worst = 0.0
    for f in range(frames):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        got = b.state()[0, :, :15]
        want = o.state()
        worst = max(worst, float(np.abs(got[:, :7] - want[:, :7]).max()))
    print("%s: worst |pose diff| over %d frames = %g" % (name, frames, worst))
    assert worst <= 1e-9
    assert not b.status().any()


@pytest.mark.parametrize("name", ["stack", "tumble", "coin", "mirror_cube", "pile", "spheres"])
def test_contact_sets_bit_exact(pkg, oracle_flavour, name):
    """Per substep: the same narrowphase pairs in the same order, the same manifold point counts, and identical
    contact points and normals (what colliders_get_contacts returns, collider.cpp:560)."""
    kw = dict(n_side=3) if name == "pile" else {}
    sc = scenes.BUILDERS[name](**kw)
    b = make(pkg, sc, n_worlds=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    o.log_enable(True)
    total = 0
    for f in range(90 if name == "coin" else 45):
        calls, contacts = b.step_logged(world=1, substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        oc, ok = o.log_get()
        o.log_clear()
        assert np.array_equal(calls, oc), (name, f)
        assert np.array_equal(contacts, ok), (name, f)
        total += len(ok)
        assert np.array_equal(b.state()[1, :, :15], o.state())
    assert total > 0


def test_golden_fixtures(pkg):
    """Committed outputs of the compiled reference (tests/golden/make_golden.py)."""
    for name in ["stack", "w256", "coin", "cube_storm"]:
        sc = scenes.BUILDERS[name]()
        b = make(pkg, sc)
        frames = sorted(int(k.split("/")[2]) for k in GOLD.files if k.startswith(name + "/state/"))
        done = 0
        for f in frames:
            while done < f:
                step(b, sc)
                done += 1
            assert np.array_equal(b.state()[0, :, :15], GOLD["%s/state/%d" % (name, f)]), (name, f)
            assert np.array_equal(b.state()[0, :, 15:21], GOLD["%s/prev_vel/%d" % (name, f)]), (name, f)


def window_digest(calls, contacts):
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(calls, dtype=np.uint32).tobytes())
    h.update(np.ascontiguousarray(contacts, dtype=np.float64).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


@pytest.mark.parametrize("name,kw,frames", [("w256", {}, 60), ("brick_wall_32x32", dict(rows=32, cols=32), 30)])
def test_timed_window_bit_exact(pkg, oracle_flavour, name, kw, frames):
    """What bench.py TIMES is what is checked: the whole W256 window (frames 0..59; the driver's K = 20 run times frames
    40..59, where a world holds ~310 contacts per substep) and 30 frames of the 32x32 brick wall as one scene, frame by
    frame against the oracle stepped alongside AND the compiled reference's committed outputs (tests/golden/windows.npz):
    state every frame, narrowphase call / contact counts every frame (pbd.cpp:584-611), and every call row and contact point
    of the last frame."""
    sc = scenes.BUILDERS[name.split("_32")[0]](**kw)
    b = make(pkg, sc, n_worlds=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    want = WINDOWS[name + "/counts"]
    c0 = b.counters()
    for f in range(frames):
        last = f + 1 == frames
        if last:
            calls, contacts = b.step_logged(world=1, substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions, max_calls=1 << 18)
        else:
            step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        got = b.state()
        assert np.array_equal(got[1, :, :15], o.state()), (name, f, np.abs(got[1, :, :15] - o.state()).max())
        assert np.array_equal(got[0], got[1])
        c1 = b.counters()
        assert ((c1["pair_tests"] - c0["pair_tests"]) // 2, (c1["contacts"] - c0["contacts"]) // 2) == tuple(want[f]), (name, f)
        c0 = c1
        key = "%s/state/%d" % (name, f + 1)
        if key in WINDOWS.files:
            assert np.array_equal(got[0, :, :15], WINDOWS[key]), key
    assert np.array_equal(window_digest(calls, contacts), WINDOWS[name + "/last_log_digest"])
    sub = refdrv.split_substeps(calls)
    assert np.array_equal(np.asarray(sub[-1], dtype=np.uint32), WINDOWS[name + "/last_substep_calls"])
    assert np.array_equal(contacts[int(sub[-1][0][3]):], WINDOWS[name + "/last_substep_contacts"])
    assert not b.status().any()


def test_known_answer_vector(pkg):
    """top cube of the reference's stack scene after 60 frames."""
    sc = scenes.stack()
    b = make(pkg, sc)
    for _ in range(60):
        step(b, sc)
    s = b.state()[0]
    assert tuple(s[8, :3]) == (0.06548885409593784, 14.041351769773531, 0.029544658022551778)


def test_sleeping_and_wakeup(pkg, oracle_flavour):
    """Island sleep bookkeeping (pbd.cpp:476-533): 600 frames of the stack scene, everything asleep at the end."""
    sc = scenes.stack()
    b = make(pkg, sc)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(600):
        step(b, sc)
        o.step()
    got, want = b.state()[0, :, :15], o.state()
    assert np.array_equal(got, want)
    assert not got[1:, 13].any()  # all cubes inactive


def test_worlds_with_different_states(pkg, oracle_flavour):
    """Independent worlds: each world of a batch, started from its own state, equals a single-world oracle run."""
    sc = scenes.tumble()
    W = 5
    b = make(pkg, sc, n_worlds=W)
    base = b.state()[0]
    rng = np.random.RandomState(11)
    states = np.repeat(base[None], W, axis=0)
    for w in range(W):
        states[w, 1:, 0] += 0.05 * rng.randn(base.shape[0] - 1)
        states[w, 1:, 7:10] = 0.3 * rng.randn(base.shape[0] - 1, 3)
    b.upload(states)
    assert np.array_equal(b.state(), states)  # upload/download round trip
    for _ in range(40):
        step(b, sc)
    got = b.state()
    for w in range(W):
        o = refdrv.RefWorld(oracle_flavour).load(sc)
        o.set_state(states[w, :, :15])
        for _ in range(40):
            o.step()
        assert np.array_equal(got[w, :, :15], o.state()), w


def test_step_host_equals_step(pkg):
    sc = scenes.cube_storm()
    a = make(pkg, sc, n_worlds=4)
    b = make(pkg, sc, n_worlds=4)
    buf_in = a.state()
    buf_out = np.zeros_like(buf_in)
    for _ in range(12):
        step(a, sc)
        b.step_host(buf_in.ctypes.data, buf_out.ctypes.data)
        buf_in, buf_out = buf_out, buf_in
    assert np.array_equal(a.state(), buf_in)


def test_broadphase_pairs(pkg, oracle_flavour):
    sc = scenes.pile(n_side=3, spacing=2.3)
    b = make(pkg, sc)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for _ in range(20):
        step(b, sc)
        o.step()
    assert np.array_equal(b.broad_pairs(0), o.broad_pairs())


def test_work_counters_match_oracle(pkg, oracle_flavour):
    """The device does the same number of narrowphase tests as the reference makes colliders_get_contacts calls."""
    sc = scenes.w256()
    b = make(pkg, sc)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    o.log_enable(True)
    for _ in range(3):
        step(b, sc)
        o.step()
    b.sync()
    calls, contacts = o.log_get()
    c = b.counters()
    assert c["pair_tests"] == len(calls)
    assert c["contacts"] == len(contacts)
    assert c["frames"] == 3
    assert np.array_equal(b.state()[0, :, :15], o.state())


def test_capacity_overflow_is_reported_not_fatal(pkg):
    """A world that runs out of contact (or pair) capacity drops work: the status word says so, and every synchronising call
    returns RP_ERR_CAPACITY -- after doing its work -- until the status is cleared."""
    sc = scenes.cube_storm()
    b = make(pkg, sc, n_worlds=2, max_contacts=8)
    for _ in range(90):
        step(b, sc)
    with pytest.raises(pkg.RawPhysCapacityError):
        b.sync()
    with pytest.raises(pkg.RawPhysCapacityError):
        b.state()
    buf = np.zeros((2, b.NB, 21))
    with pytest.raises(pkg.RawPhysCapacityError):
        b.step_host(None, buf.ctypes.data)
    st = b.status()
    assert st[0] & 64 and st[1] & 64  # RP_ST_CONTACT_CAPACITY
    assert np.isfinite(buf).all() and np.isfinite(b.state(ignore_capacity=True)).all()
    b.clear_status()
    b.sync()
    # too few pairs
    b = make(pkg, sc, max_pairs=4)
    step(b, sc)
    with pytest.raises(pkg.RawPhysCapacityError):
        b.sync()
    assert b.status()[0] & 128  # RP_ST_PAIR_CAPACITY


def test_state_transfer_ranges_are_checked(pkg):
    sc = scenes.stack()
    b = make(pkg, sc, n_worlds=3)
    buf = np.zeros((4, b.NB, 21))
    L = pkg.lib()
    for first, n in [(0xFFFFFFFF, 2), (3, 1), (2, 2), (0, 0), (0, 4)]:
        assert L.rp_batch_download_state(b.h, first, n, buf.ctypes.data) == 1
        assert L.rp_batch_upload_state(b.h, first, n, buf.ctypes.data) == 1
    assert L.rp_batch_download_state(b.h, 2, 1, buf.ctypes.data) == 0


def test_large_batch_invariants(pkg):
    """BASELINE size (4096 worlds x 257 bodies): size-independent properties -- every world of a uniform batch stays
    bit-identical to world 0, world 0 equals the committed fixture, no status bits."""
    sc = scenes.w256()
    b = make(pkg, sc, n_worlds=4096)
    for _ in range(5):
        step(b, sc)
    got = b.state()
    assert np.array_equal(got[0, :, :15], GOLD["w256/state/5"])
    assert (got == got[0][None]).all()
    assert not b.status().any()


@pytest.mark.parametrize("name", ["w256", "pile", "tumble", "coin", "mirror_cube", "cube_storm"])
def test_bounds_cull_changes_nothing(pkg, name):
    """k_cull's bounds test must be invisible: with and without it the batch evolves bit-identically (and the number of
    narrowphase pair visits, which the cull does not change, stays the reference's)."""
    kw = dict(n_side=3) if name == "pile" else {}
    sc = scenes.BUILDERS[name](**kw)
    frames = 25 if name == "w256" else 90
    a = make(pkg, sc, n_worlds=2)
    b = make(pkg, sc, n_worlds=2, disable_cull=True)
    for _ in range(frames):
        step(a, sc)
        step(b, sc)
    assert np.array_equal(a.state(), b.state())
    ca, cb = a.counters(), b.counters()
    assert ca["pair_tests"] == cb["pair_tests"] and ca["contacts"] == cb["contacts"] and ca["gjk_hits"] == cb["gjk_hits"]


def test_brick_wall_32x32_single_scene(pkg, oracle_flavour):
    """Config 3: one large scene (1025 bodies), deep dependency chains; level-major solve vs the sequential oracle."""
    sc = scenes.brick_wall(rows=32, cols=32)
    b = make(pkg, sc)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(3):
        step(b, sc)
        o.step()
        assert np.array_equal(b.state()[0, :, :15], o.state()), f
    assert not b.status().any()


def test_config2_4096_copies_of_the_stack_scene(pkg, oracle_flavour):
    """BASELINE config 2: 4096 copies of the reference's stack scene on one GPU for 600 frames; every
    world bitwise equal to world 0, and world 0 equal to the oracle at frames 1, 60 and 600 (all asleep by then)."""
    sc = scenes.stack()
    b = make(pkg, sc, n_worlds=4096)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(1, 601):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if f in (1, 60, 600):
            got = b.state()
            assert np.array_equal(got[0, :, :15], o.state()), f
            assert (got == got[0]).all(), f
    assert not got[0, 1:, 13].any()  # every cube asleep
    assert not b.status().any()


def test_config5_joint_worlds_sharded(pkg, oracle_flavour):
    """BASELINE config 5 in small: hinge-lever worlds split over two batches the way multi.partition splits them over
    ranks; each shard's worlds equal the oracle's single world within the joint tolerance."""
    import __graft_entry__ as ge
    ge.load_package()
    from rawphys_b200 import multi
    sc = scenes.hinge_joints()
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    shards = [make(pkg, sc, n_worlds=n) for (_, n) in multi.partition(257, 2)]
    assert [s.W for s in shards] == [129, 128]
    for f in range(40):
        for s in shards:
            step(s, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    want = o.state()
    for s in shards:
        got = s.state()
        assert (got == got[0]).all()
        assert np.abs(got[0, :, :7] - want[:, :7]).max() <= 1e-9


@pytest.mark.parametrize("name,substeps,iters,dt", [("stack", 5, 3, 1.0 / 60.0), ("tumble", 8, 2, 1.0 / 90.0), ("hinge_joints", 10, 4, 1.0 / 60.0)])
def test_other_step_settings(pkg, oracle_flavour, name, substeps, iters, dt):
    """pbd_simulate's remaining arguments (pbd.cpp:464-472): dt, num_substeps and num_pos_iters other than the examples'
    1/60, 20, 1 -- several positional sweeps per substep go through the same level schedule."""
    sc = scenes.BUILDERS[name]()
    b = make(pkg, sc, n_worlds=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(40):
        b.step(dt=dt, substeps=substeps, iters=iters, collisions=sc.collisions)
        o.step(dt=dt, substeps=substeps, iters=iters, collisions=sc.collisions)
    got, want = b.state()[0, :, :15], o.state()
    if sc.constraints:
        assert np.abs(got[:, :7] - want[:, :7]).max() <= 1e-9
    else:
        assert np.array_equal(got, want)
    assert not b.status().any()


def test_collisions_disabled_and_zero_dt(pkg, oracle_flavour):
    """enable_collisions = false skips broadphase-to-contact work entirely (pbd.cpp:584); dt <= 0 is a no-op (:471)."""
    sc = scenes.stack()
    b = make(pkg, sc, n_worlds=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    before = b.state()
    b.step(dt=0.0)
    b.step(dt=-1.0)
    assert np.array_equal(b.state(), before)
    for f in range(20):
        b.step(collisions=False)
        o.step(collisions=False)
    assert np.array_equal(b.state()[0, :, :15], o.state())
    assert b.counters()["pair_tests"] == 0


@pytest.mark.parametrize("n_worlds", [1, 31, 33])
def test_ragged_world_counts(pkg, oracle_flavour, n_worlds):
    """world counts that do not fill a warp's 32 lanes: the padded lanes must neither disturb the live ones nor be visible"""
    sc = scenes.stack()
    b = make(pkg, sc, n_worlds=n_worlds)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(30):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    got = b.state()
    assert got.shape[0] == n_worlds and (got == got[0]).all()
    assert np.array_equal(got[0, :, :15], o.state())
    assert b.status().shape == (n_worlds,) and not b.status().any()


def test_world_without_pairs(pkg, oracle_flavour):
    """a single free body: empty broadphase, empty schedule, nothing to solve -- ballistic motion only"""
    sc = scenes.Scene("lonely")
    sc.bodies.append(scenes.BodyDesc((0.0, 5.0, 0.0), scenes.IDENT, 2.0, False, [scenes.hull("cube", (1.0, 1.0, 1.0))]))
    b = make(pkg, sc, n_worlds=3)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(10):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    assert np.array_equal(b.state()[0, :, :15], o.state())
    c = b.counters()
    assert c["broad_pairs"] == 0 and c["contacts"] == 0 and not b.status().any()


@pytest.mark.parametrize("name", ["stack", "tumble", "coin"])
def test_previous_velocities_match(pkg, oracle_flavour, name):
    """columns 15..20 of the state record: Entity::previous_linear_velocity / previous_angular_velocity as pbd_simulate
    leaves them (pbd.cpp:629-630: the velocities each body had before the LAST substep's derivation)."""
    sc = scenes.BUILDERS[name]()
    b = make(pkg, sc, n_worlds=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(45):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if f % 11 == 0 or f == 44:
            got = b.state()[0]
            assert np.array_equal(got[:, :15], o.state()), f
            assert np.array_equal(got[:, 15:21], o.prev_velocities()), f


def test_spot_storm_compound_large_hulls(pkg, oracle_flavour):
    """spot_storm.cpp: 8 compound bodies x 11 hulls of up to 529 vertices. Collider pairs of one body
    pair stay in the reference's (i outer, j inner) order, hulls too large for the per-thread staging block go through the
    warp-per-pair kernels (k_gjk_warp / k_epa_warp; hulls above 128 vertices are scanned in place): bit-exact trajectories
    and per-substep contact sets."""
    sc = scenes.spot_storm(n=2)
    b = make(pkg, sc, n_worlds=2, max_pairs=8192, max_contacts=8192)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(30):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if f % 5 == 4:
            got = b.state()
            assert np.array_equal(got[0, :, :15], o.state()), (f, np.abs(got[0, :, :15] - o.state()).max())
            assert np.array_equal(got[0], got[1])
    assert not b.status().any()
    # one more frame with the contact log of every substep
    o.log_enable(True)
    o.log_clear()
    o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    want_calls, want_contacts = o.log_get()
    o.log_enable(False)
    calls, contacts = b.step_logged(world=1, substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    assert np.array_equal(calls, want_calls) and np.array_equal(contacts, want_contacts)
    assert want_contacts.shape[0] > 0


@pytest.mark.parametrize("name,wpb", [("stack", 1), ("tumble", 3), ("coin", 4), ("mirror_cube", 2), ("hinge_joints", 4), ("mutual_orientation", 5),
                                      ("cube_storm", 4), ("spheres", 64)])
def test_world_block_sweeps_bit_exact(pkg, oracle_flavour, name, wpb):
    """k_solve_block (one CTA per block of `wpb` worlds, both sweeps, CTA-scoped barriers between levels) is what large batches
    run; small test batches would get the level-major cooperative sweeps, so it is forced here: same per-world level order,
    hence the same bits as the sequential reference (contact scenes) / the joint tolerance (libm scenes), with world counts
    that do not divide into blocks and worlds started from different states."""
    sc = scenes.BUILDERS[name]()
    W = 7
    b = make(pkg, sc, n_worlds=W, sweep_block_worlds=wpb)
    base = b.state()[0].copy()
    states = np.repeat(base[None], W, axis=0)
    rng = np.random.RandomState(5)
    moving = [i for i, body in enumerate(sc.bodies) if not body.fixed]
    for w in range(1, W, 2):  # odd worlds start perturbed: their level lists differ from their neighbours'
        states[w, moving, 0] += 0.02 * rng.randn(len(moving))
        states[w, moving, 7:10] += 0.2 * rng.randn(len(moving), 3)
    b.upload(states)
    frames = 60
    for _ in range(frames):
        step(b, sc)
    got = b.state()
    for w in (0, 1, 4, 5):
        o = refdrv.RefWorld(oracle_flavour).load(sc)
        o.set_state(states[w, :, :15])
        for _ in range(frames):
            o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if sc.constraints and name != "mutual_orientation":
            assert np.abs(got[w, :, :7] - o.state()[:, :7]).max() <= 1e-9, (name, w)
        else:
            assert np.array_equal(got[w, :, :15], o.state()), (name, w, np.abs(got[w, :, :15] - o.state()).max())
    assert not b.status().any()


def test_world_block_sweeps_deep_levels(pkg, oracle_flavour):
    """compound bodies chain 121 units per body pair: 1300+ mostly empty levels per block"""
    sc = scenes.spot_storm(n=2)
    b = make(pkg, sc, n_worlds=3, max_pairs=8192, max_contacts=8192, sweep_block_worlds=2)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for f in range(30):
        step(b, sc)
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    got = b.state()
    assert np.array_equal(got[0, :, :15], o.state()) and np.array_equal(got[0], got[2])
    assert not b.status().any()


def test_islands_disabled_nothing_sleeps(pkg, oracle_flavour):
    """rp_batch_cfg.disable_islands = the reference compiled without ENABLE_SIMULATION_ISLANDS (pbd.cpp:12, :476-533): no
    deactivation timers, every body stays active. Until the first island of the reference falls asleep the two are the same
    computation (bit for bit); afterwards the reference's cubes rest while these keep being solved."""
    sc = scenes.stack()
    b = make(pkg, sc, n_worlds=2, disable_islands=True)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    diverged_at = None
    for f in range(600):
        step(b, sc)
        o.step()
        if diverged_at is None and not o.state()[1:, 13].all():
            diverged_at = f
        if diverged_at is None and f % 20 == 19:
            assert np.array_equal(b.state()[0, :, :13], o.state()[:, :13]), f
    st = b.state()[0]
    assert diverged_at is not None and diverged_at > 100
    assert st[:, 13].all() and not st[:, 14].any()  # all active, timers untouched
    assert not b.status().any()


def test_bodies_thrown_in_mid_run(pkg, oracle_flavour):
    """examples_util_throw_object (examples_util.cpp:52-95): an icosahedron, then a sphere, created between frames and thrown at
    the stack at 15 m/s. rp_batch_create_from carries every world's state over device to
    device; the reference creates the entities in its global table at the same frames. Bit for bit throughout."""
    sc = scenes.stack()
    b = make(pkg, sc, n_worlds=3)
    o = refdrv.RefWorld(oracle_flavour).load(sc)

    def both(frames):
        for _ in range(frames):
            b.step()
            o.step()
        assert np.array_equal(b.state()[0, :, :15], o.state())
        assert np.array_equal(b.state()[0], b.state()[2])

    both(25)
    thrown = [scenes.BodyDesc((6.0, 4.0, 0.5), scenes.quaternion_new((0.35, 0.44, 0.12), 0.0), 1.0, False, [scenes.hull("ico", (1.0, 1.0, 1.0))], 0.8, 0.8, 0.0),
              scenes.BodyDesc((-5.0, 9.0, 0.2), scenes.quaternion_new((0.35, 0.44, 0.12), 0.0), 1.0, False, [scenes.sphere(1.0)], 0.8, 0.8, 0.0)]
    vel = [(-15.0, 0.0, 0.0), (12.0, -3.0, 0.0)]
    for k in range(2):
        sc.bodies.append(thrown[k])
        nb = len(sc.bodies)
        nbatch = pkg.Batch.create_from(pkg.Scene(sc), b, list(range(nb - 1)) + [-1])
        nbatch.set_scene_forces(sc)
        st = nbatch.state()
        st[:, nb - 1, 7:10] = vel[k]  # e->linear_velocity = ... (examples_util.cpp:94)
        nbatch.upload(st)
        b.close()
        b = nbatch
        o.add_body(thrown[k])
        ost = o.state()
        ost[nb - 1, 7:10] = vel[k]
        o.set_state(ost)
        both(20)
    assert np.abs(b.state()[0, 1:9, 7:13]).max() > 0.5  # the stack was hit
    assert not b.status().any()
