"""CPU tests of the product's host side (-m "not gpu"): the C-ABI library loads, exports every symbol the header
declares, builds scene templates identical to the reference's, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import refdrv
import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HULLS = np.load(os.path.join(ROOT, "tests", "golden", "hulls.npz"))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "rawphys_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    L = pkg.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(pkg.EXPORTS) == declared


def test_product_does_not_touch_the_oracle():
    """The product (package + header) must never import, link or execute anything under oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "raw-physics_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "libport_oracle" not in text and "libref_oracle" not in text, f
                assert "import refdrv" not in text and "oracle/" not in text, f


@pytest.mark.parametrize("name", ["stack", "brick_wall", "coin", "pile", "mirror_cube", "spheres", "arm"])
def test_scene_params_match_reference(pkg, name):
    """entity_create_ex results (entity.cpp:24-65): inverse mass, inertia (quirk q9), inverse inertia, bounding radius."""
    kw = dict(n_side=3) if name == "pile" else {}
    sc = scenes.BUILDERS[name](**kw)
    got = pkg.Scene(sc).params()
    assert np.array_equal(got, GOLD[name + "/params"])


@pytest.mark.parametrize("mesh", ["cube", "floor", "ico", "ramp", "cylinder", "lever", "seesaw_support"])
def test_scene_hulls_match_reference(pkg, mesh):
    sc = scenes.Scene("h")
    sc.bodies.append(scenes.BodyDesc((0, 0, 0), scenes.IDENT, 1.0, False, [scenes.hull(mesh, (1.0, 1.0, 1.0))]))
    h = pkg.Scene(sc).hull(0)
    for k, v in h.items():
        assert np.array_equal(v, HULLS["%s/%s" % (mesh, k)]), (mesh, k)


def test_sphere_collider_has_no_hull(pkg):
    assert pkg.Scene(scenes.spheres(2)).hull(1) is None


def test_bad_arguments_are_refused(pkg):
    L = pkg.lib()
    s = pkg.Scene()
    v = np.zeros((3, 3))
    idx = np.array([0, 1, 7], dtype=np.uint32)  # index out of range
    assert L.rp_scene_collider_hull(s.h, v.ctypes.data_as(C.POINTER(C.c_double)), 3, idx.ctypes.data_as(C.POINTER(C.c_uint32)), 3) == -1
    assert L.rp_scene_collider_hull(s.h, None, 3, None, 3) == -1
    assert L.rp_scene_add_mutual_orientation_constraint(s.h, 0, 1, 0.0) == -1  # no such bodies
    out = C.c_void_p()
    assert L.rp_batch_create(s.h, 1, 0, None, C.byref(out)) == 1  # RP_ERR_ARG: empty scene
    assert L.rp_batch_create(None, 1, 0, None, C.byref(out)) == 1
    assert L.rp_batch_step(None, 1.0 / 60, 20, 1, 1) == 1


def test_no_cpu_fallback(pkg):
    """Without a CUDA device batch creation fails loudly with RP_ERR_CUDA; nothing is computed on the host."""
    L = pkg.lib()
    if L.rp_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.RawPhysError) as e:
        pkg.Batch(pkg.Scene(scenes.stack()), n_worlds=2)
    assert "code 2" in str(e.value)


def test_build_flags_forbid_fma():
    """Parity depends on --fmad=false (SURVEY.md TL;DR 3) and on sm_100a being the only target."""
    text = open(os.path.join(ROOT, "raw-physics_b200", "build.py")).read()
    assert "--fmad=false" in text and "arch=compute_100a,code=sm_100a" in text and "-lineinfo" in text
    assert "use_fast_math" not in text


def test_adopted_topology_and_params_round_trip(pkg):
    """rp_scene_collider_hull_topology / rp_scene_add_body_params (what the pbd_simulate shim feeds from the reference's
    own Collider_Convex_Hull / Entity structs) reproduce the template built from the triangle soups, array for array."""
    import ctypes as C
    sc = scenes.BUILDERS["coin"]()
    built = pkg.Scene(sc)
    params = built.params()
    L = pkg.lib()
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    adopted = pkg.Scene()
    for i, b in enumerate(sc.bodies):
        for c in range(len(b.colliders)):
            h = built.hull(i, c)
            if h is None:
                assert L.rp_scene_collider_sphere(adopted.h, C.c_float(b.colliders[c].radius)) >= 0
                continue
            args = [h[k].ctypes.data_as(up) for k in ("face_ptr", "face_idx", "v2f_ptr", "v2f_idx", "v2n_ptr", "v2n_idx", "f2n_ptr", "f2n_idx")]
            assert L.rp_scene_collider_hull_topology(adopted.h, h["verts"].ctypes.data_as(dp), h["verts"].shape[0],
                                                     h["normals"].ctypes.data_as(dp), h["normals"].shape[0], *args) >= 0
        p = params[i]
        pos = np.asarray(b.position, dtype=np.float64)
        rot = np.asarray(b.rotation, dtype=np.float64)
        inertia, inv = np.ascontiguousarray(p[1:10]), np.ascontiguousarray(p[10:19])
        assert L.rp_scene_add_body_params(adopted.h, pos.ctypes.data_as(dp), rot.ctypes.data_as(dp), p[0], inertia.ctypes.data_as(dp),
                                          inv.ctypes.data_as(dp), p[19], int(p[23]), p[20], p[21], p[22]) == i
    assert np.array_equal(adopted.params(), params)
    for i, b in enumerate(sc.bodies):
        for c in range(len(b.colliders)):
            h0, h1 = built.hull(i, c), adopted.hull(i, c)
            assert (h0 is None) == (h1 is None)
            if h0 is not None:
                assert all(np.array_equal(h0[k], h1[k]) for k in h0)
    # malformed CSR input is refused, not adopted
    h = built.hull(1, 0)
    bad = h["face_idx"].copy()
    bad[0] = 10 ** 6
    args = [h[k].ctypes.data_as(up) for k in ("face_ptr",)] + [bad.ctypes.data_as(up)] + [
        h[k].ctypes.data_as(up) for k in ("v2f_ptr", "v2f_idx", "v2n_ptr", "v2n_idx", "f2n_ptr", "f2n_idx")]
    assert L.rp_scene_collider_hull_topology(adopted.h, h["verts"].ctypes.data_as(dp), h["verts"].shape[0],
                                             h["normals"].ctypes.data_as(dp), h["normals"].shape[0], *args) == -1


def test_shim_build_links_against_the_library(pkg):
    """oracle/_ref/libref_shim.so (the unmodified reference + raw-physics_b200/shim/pbd_b200.cpp, `make -C oracle shim`)
    loads, resolves librawphys_b200.so through its rpath and builds scenes with the reference's own code; stepping it
    needs a GPU (tests/test_gpu_shim.py)."""
    import refdrv
    if not (refdrv.available("shim") and refdrv.available("strict")):
        pytest.skip("built only where /root/reference is present")
    sc = scenes.BUILDERS["mirror_cube"]()
    a, b = refdrv.RefWorld("strict").load(sc), refdrv.RefWorld("shim").load(sc)
    assert np.array_equal(a.params(), b.params()) and np.array_equal(a.state(), b.state())
    src = open(os.path.join(ROOT, "raw-physics_b200", "shim", "pbd_b200.cpp")).read()
    for needle in ("rp_scene_collider_hull_topology", "rp_scene_add_body_params", "rp_batch_step_host", "_Z12pbd_simulatedPP6Entityjji",
                   "_Z29pbd_simulate_with_constraintsdPP6EntityP10Constraintjji"):
        assert needle in src


def test_scene_builder_validates_its_inputs(pkg):
    """The header's contract: a status code instead of the reference's assert / silent garbage."""
    L = pkg.lib()
    s = L.rp_scene_create()
    cube = scenes.hull("cube", (1.0, 1.0, 1.0))
    v = np.ascontiguousarray(cube.vertices)
    idx = np.ascontiguousarray(cube.indices)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    pos, quat = np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0])
    d = lambda a: a.ctypes.data_as(dp)
    # a non-fixed body needs a positive mass ...
    assert L.rp_scene_collider_hull(s, d(v), v.shape[0], idx.ctypes.data_as(up), idx.shape[0]) >= 0
    assert L.rp_scene_add_body(s, d(pos), d(quat), 0.0, 0, 0.5, 0.5, 0.0) == -1
    assert L.rp_scene_add_body(s, d(pos), d(quat), -1.0, 0, 0.5, 0.5, 0.0) == -1
    assert L.rp_scene_add_body(s, d(pos), d(quat), 1.0, 0, 0.5, 0.5, 0.0) == 0
    # ... and an invertible inertia tensor: a body without any collider has none
    assert L.rp_scene_add_body(s, d(pos), d(quat), 1.0, 0, 0.5, 0.5, 0.0) == -1
    assert L.rp_scene_add_body(s, d(pos), d(quat), 0.0, 1, 0.5, 0.5, 0.0) == 1  # (fixed: fine)
    # zero-area triangles have no normal
    flat = np.array([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]])
    tri = np.arange(3, dtype=np.uint32)
    assert L.rp_scene_collider_hull(s, d(flat), 3, tri.ctypes.data_as(up), 3) == -1
    # constraints: NULL vectors, axis selectors out of range
    r = np.zeros(3)
    assert L.rp_scene_add_positional_constraint(s, 0, 1, None, d(r), 0.0, d(r)) == -1
    assert L.rp_scene_add_positional_constraint(s, 0, 1, d(r), d(r), 0.0, None) == -1
    assert L.rp_scene_add_hinge_joint_constraint(s, 0, 1, d(r), None, 0.0, 0, 0, 0, 0, 0, 0.0, 0.0) == -1
    assert L.rp_scene_add_hinge_joint_constraint(s, 0, 1, d(r), d(r), 0.0, 6, 0, 0, 0, 0, 0.0, 0.0) == -1
    assert L.rp_scene_add_spherical_joint_constraint(s, 0, 1, d(r), d(r), 0, 0, -1, 0, 0.0, 0.0, 0.0, 0.0) == -1
    assert L.rp_scene_add_positional_constraint(s, 0, 1, d(r), d(r), 0.0, d(r)) == 0
    # output pointers
    assert L.rp_scene_hull_sizes(s, 0, 0, None) != 0
    sizes = np.zeros(6, dtype=np.int32)
    assert L.rp_scene_hull_sizes(s, 0, 0, sizes.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    assert L.rp_scene_hull_dump(s, 0, 0, None, None, None, None, None, None, None, None, None, None) != 0
    L.rp_scene_destroy(s)


def test_device_hull_builder_needs_a_device(pkg):
    """rp_scene_set_hull_device: -1 (host) always works; a CUDA device that does not exist is an error, not a silent host build"""
    L = pkg.lib()
    s = pkg.Scene()
    assert L.rp_scene_set_hull_device(s.h, -1) == 0
    assert L.rp_scene_set_hull_device(s.h, 4096) != 0
    assert s.hull_build_stats() == (0, 0.0)
