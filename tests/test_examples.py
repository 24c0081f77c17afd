"""The built-in scenes of the library (raw-physics_b200/csrc/rp_examples.cpp: the init() halves of the reference's 14 examples
and the benchmark worlds, host C++) against the test-side descriptions (tests/scenes.py, which feed the oracle): every
body, collider soup, constraint, initial velocity and derived parameter bit for bit. CPU only -- scene construction needs no GPU.
"""
import numpy as np
import pytest

import scenes

CASES = {"stack": ((), {}), "w256": ((), {}), "brick_wall": ((32, 32), dict(rows=32, cols=32)), "cube_storm": ((), {}), "seesaw": ((), {}),
         "cube_and_ramp": ((), {}), "coin": ((), {}), "spring": ((), {}), "hinge_joints": ((), {}), "arm": ((), {}), "triple_pendula": ((), {}),
         "rott_pendulum": ((), {}), "mirror_cube": ((), {}), "spot_storm": ((2, 99), dict(n=2)), "pile": ((3, 12345, 2.3), dict(n_side=3, spacing=2.3)),
         "tumble": ((), {}), "spheres": ((), {})}
REFERENCE_EXAMPLES = ["arm", "brick_wall", "coin", "cube_and_ramp", "cube_storm", "debug", "hinge_joints", "mirror_cube", "rott_pendulum", "seesaw",
                      "spot_storm", "spring", "stack", "triple_pendula"]  # src/examples/*.cpp minus examples_util


def test_every_reference_example_is_built_in(pkg):
    names = pkg.example_names()
    assert all(n in names for n in REFERENCE_EXAMPLES)
    with pytest.raises(pkg.RawPhysError):
        pkg.example("no_such_scene")


@pytest.mark.parametrize("name", sorted(CASES))
def test_example_equals_test_scene(pkg, name):
    params, kw = CASES[name]
    sc, desc = pkg.example(name, params, perturb=True)
    ref = scenes.BUILDERS[name](**kw)
    assert len(desc.bodies) == len(ref.bodies)
    for a, b in zip(desc.bodies, ref.bodies):
        assert tuple(a.position) == tuple(map(float, b.position)) and tuple(a.rotation) == tuple(map(float, b.rotation))
        assert (a.mass, a.fixed, a.mu_s, a.mu_d, a.restitution) == (b.mass, b.fixed, b.mu_s, b.mu_d, b.restitution)
        assert len(a.colliders) == len(b.colliders)
        for ca, cb in zip(a.colliders, b.colliders):
            assert ca.kind == cb.kind
            if ca.kind == "sphere":
                assert ca.radius == cb.radius
            else:
                assert np.array_equal(ca.vertices, cb.vertices) and np.array_equal(ca.indices, cb.indices)
    assert len(desc.constraints) == len(ref.constraints)
    for a, b in zip(desc.constraints, ref.constraints):
        for k, v in b.items():
            if isinstance(v, (tuple, list, np.ndarray)):
                assert tuple(map(float, v)) == tuple(a[k]), (name, k)
            else:
                assert a[k] == v, (name, k)
    assert (desc.initial_state is None) == (ref.initial_state is None)
    if ref.initial_state is not None:
        assert np.array_equal(desc.initial_state, ref.initial_state)
    assert np.array_equal(sc.params(), pkg.Scene(ref).params())
    if name != "spot_storm":  # spot_storm.cpp:184 steps with ONE substep; the test scene keeps 20
        assert (desc.substeps, desc.iters, desc.collisions) == (ref.substeps, ref.iters, ref.collisions)
    else:
        assert (desc.substeps, desc.iters, desc.collisions) == (1, 1, True)


def test_unperturbed_joint_examples_start_at_rest(pkg):
    for name in ["hinge_joints", "arm", "triple_pendula", "rott_pendulum"]:
        sc, desc = pkg.example(name)
        assert desc.initial_state is None and not sc.initial_state()[:, 7:13].any()


def test_debug_example_is_the_floor_alone(pkg):
    """debug.cpp:47-66: floor.obj at (0, -2, 0), fixed; the user spawns everything else"""
    sc, desc = pkg.example("debug")
    assert len(desc.bodies) == 1 and desc.bodies[0].fixed and tuple(desc.bodies[0].position) == (0.0, -2.0, 0.0)
    want = scenes.hull("floor", (1.0, 1.0, 1.0)).vertices
    assert np.array_equal(desc.bodies[0].colliders[0].vertices, want)
