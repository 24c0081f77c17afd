"""Worker of tests/test_gpu_f32.py: runs scenes with the single-precision build (RAWPHYS_B200_LIB = librawphys_b200_f32.so) in a
process of its own and saves states / status words.   python tests/f32_worker.py OUT.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
assert pkg.LIB_PATH.endswith("_f32.so"), pkg.LIB_PATH
out = {}


def run(tag, name, params=(), perturb=False, worlds=3, frames=(60,), **kw):
    scene, desc = pkg.example(name, params, perturb=perturb)
    b = pkg.Batch(scene, n_worlds=worlds, device=0, **kw)
    b.set_scene_forces(desc)
    for f in range(1, max(frames) + 1):
        b.step(1.0 / 60.0, desc.substeps, desc.iters, desc.collisions)
        if f in frames:
            st = b.state()
            out["%s/state/%d" % (tag, f)] = st[worlds - 1]
            out["%s/same/%d" % (tag, f)] = np.array([np.array_equal(st[0], st[worlds - 1])])
    out["%s/status" % tag] = b.status()
    out["%s/fixed" % tag] = np.array([bool(bd.fixed) for bd in desc.bodies])
    b.close()


run("stack", "stack", frames=(10, 60, 240))
run("stack70", "stack", worlds=70, frames=(60,))           # dataflow sweeps
run("w256", "w256", (2, 2, 4), worlds=33, frames=(60, 120))
run("wall", "brick_wall", (8, 8), worlds=1, frames=(90,), coloured=True)
run("levers", "hinge_joints", perturb=True, worlds=4, frames=(60,))
run("coin", "coin", worlds=2, frames=(60,))
run("spheres", "spheres", worlds=2, frames=(120,))
np.savez(sys.argv[1], **out)
