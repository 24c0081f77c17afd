"""Drop-in check of the boundary (-m gpu): the reference's OWN object code -- entity table, collider construction,
constraint init, force bookkeeping, the update() sequence of its examples (oracle/ref_driver.cpp) -- with nothing but its
two pbd_simulate entry points (src/physics/pbd.h:93-94) redirected to the product's same-signature shim
raw-physics_b200/shim/pbd_b200.cpp, which gathers the reference's Entity / Collider_Convex_Hull / Constraint structs and
steps on the GPU through the C ABI. oracle/_ref/libref_shim.so is built by `make -C oracle shim` in the build container.

The world stepped through the shim must equal the all-reference world bit for bit (contact scenes) or within the libm
tolerance of tests/test_gpu_parity.py (joint-limit scenes).
"""
import numpy as np
import pytest

import refdrv
import scenes

pytestmark = pytest.mark.gpu

needs_shim = pytest.mark.skipif(not (refdrv.available("shim") and refdrv.available("strict")),
                                reason="oracle/_ref/libref_shim.so is built only where /root/reference is present")


def run_both(name, frames, **kw):
    sc = scenes.BUILDERS[name](**kw)
    ref = refdrv.RefWorld("strict").load(sc)
    shim = refdrv.RefWorld("shim").load(sc)
    worst = 0.0
    for f in range(frames):
        ref.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        shim.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        a, b = ref.state(), shim.state()
        worst = max(worst, float(np.abs(a - b).max()))
        yield f, a, b, worst


@needs_shim
@pytest.mark.parametrize("name,frames", [("stack", 90), ("cube_and_ramp", 60), ("spheres", 60), ("mirror_cube", 60), ("coin", 45)])
def test_reference_scene_through_shim_bit_exact(pkg, name, frames):
    for f, a, b, worst in run_both(name, frames):
        assert np.array_equal(a, b), (name, f, worst)


@needs_shim
def test_sleeping_through_shim(pkg):
    """600 frames of the stack scene: every cube asleep at the end, identical sleep bookkeeping (pbd.cpp:476-533)."""
    last = None
    for f, a, b, worst in run_both("stack", 600):
        if f % 50 == 49:
            assert np.array_equal(a, b), (f, worst)
        last = b
    assert not last[1:, 13].any()


@needs_shim
@pytest.mark.parametrize("name,frames", [("hinge_joints", 60), ("arm", 60)])
def test_joint_scene_through_shim(pkg, name, frames):
    worst = 0.0
    for f, a, b, worst in run_both(name, frames):
        pass
    print("%s through the shim: worst |state diff| over %d frames = %g" % (name, frames, worst))
    assert worst <= 1e-9


@needs_shim
def test_shim_follows_scene_changes(pkg):
    """The entity table changes between calls (a body is added, as examples_util_throw_object does,
    examples_util.cpp:52-95): the shim rebuilds its template and carries on from the entities' current state."""
    sc = scenes.BUILDERS["stack"]()
    worlds = [refdrv.RefWorld("strict").load(sc), refdrv.RefWorld("shim").load(sc)]
    for f in range(20):
        for w in worlds:
            w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
    cube = sc.bodies[1].colliders[0]
    for w in worlds:
        L = w.lib
        L.ref_collider_begin()
        v = np.ascontiguousarray(cube.vertices, dtype=np.float64)
        idx = np.ascontiguousarray(cube.indices, dtype=np.uint32)
        L.ref_collider_add_hull(refdrv._d(v), v.shape[0], refdrv._u(idx), idx.shape[0])
        pos = np.array([0.3, 25.0, 0.1])
        quat = np.array([0.0, 0.0, 0.0, 1.0])
        L.ref_entity_create(refdrv._d(pos), refdrv._d(quat), 2.0, 0, 0.6, 0.5, 0.0)
    for f in range(60):
        for w in worlds:
            w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        a, b = worlds[0].state(), worlds[1].state()
        assert a.shape[0] == len(sc.bodies) + 1
        assert np.array_equal(a, b), (f, np.abs(a - b).max())
