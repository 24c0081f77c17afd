"""Device-side hull construction (csrc/rp_hull.cuh) against the reference's collider_convex_hull_create
(collider.cpp:194-364): the golden topology of the compiled reference for 7 meshes, the host build for every mesh the
library ships (the 11 spot hulls up to 529 vertices included), and a trajectory stepped from device-built hulls."""
import os

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HULLS = np.load(os.path.join(ROOT, "tests", "golden", "hulls.npz"))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
MESHES = sorted(f[:-4] for f in os.listdir(os.path.join(ROOT, "tests", "golden", "meshes")) if f.endswith(".f32"))


def one_hull_scene(mesh, scale=(1.0, 1.0, 1.0)):
    sc = scenes.Scene("h")
    sc.bodies.append(scenes.BodyDesc((0, 0, 0), scenes.IDENT, 1.0, False, [scenes.hull(mesh, scale)]))
    return sc


@pytest.mark.parametrize("mesh", ["cube", "floor", "ico", "ramp", "cylinder", "lever", "seesaw_support"])
def test_device_hull_matches_reference_fixture(pkg, mesh):
    h = pkg.Scene(one_hull_scene(mesh), hull_device=0).hull(0)
    for k, v in h.items():
        assert np.array_equal(v, HULLS["%s/%s" % (mesh, k)]), (mesh, k)


@pytest.mark.parametrize("mesh", MESHES)
def test_device_hull_equals_host_build(pkg, mesh):
    """every array of the topology, bit for bit (normals as bit patterns), for every shipped mesh at a non-trivial scale"""
    sc = one_hull_scene(mesh, (1.5, 0.75, 2.0))
    host, dev = pkg.Scene(sc), pkg.Scene(sc, hull_device=0)
    a, b = host.hull(0), dev.hull(0)
    assert sorted(a) == sorted(b)
    for k in a:
        if a[k].dtype == np.float64:
            assert np.array_equal(a[k].view(np.uint64), b[k].view(np.uint64)), (mesh, k)
        else:
            assert np.array_equal(a[k], b[k]), (mesh, k)
    assert dev.hull_build_stats()[0] == 1
    assert np.array_equal(host.params(), dev.params())


def test_degenerate_soup_is_refused_by_the_device_build(pkg):
    import ctypes as C
    L = pkg.lib()
    s = pkg.Scene(hull_device=0)
    flat = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0]], dtype=np.float64)  # zero-area triangle
    tri = np.array([0, 1, 2], dtype=np.uint32)
    assert L.rp_scene_collider_hull(s.h, flat.ctypes.data_as(C.POINTER(C.c_double)), 3, tri.ctypes.data_as(C.POINTER(C.c_uint32)), 3) == -1
    assert L.rp_scene_set_hull_device(s.h, 99) != 0


def test_stack_from_device_built_hulls_is_bit_exact(pkg):
    """the reference's stack scene with its hulls built on the GPU, 10 frames against the compiled reference's fixture"""
    sc, desc = pkg.example("stack", hull_device=0)
    assert sc.hull_build_stats()[0] >= 2
    batch = pkg.Batch(sc, n_worlds=2, device=0)
    batch.set_scene_forces(desc)
    for f in range(1, 11):
        batch.step()
        key = "stack/state/%d" % f
        if key in GOLD:
            assert np.array_equal(batch.state()[1, :, :13], GOLD[key][:, :13]), f
    batch.close()
