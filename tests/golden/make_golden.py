"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref/libref_oracle.so
(oracle/Makefile `ref`; needs /root/reference, i.e. the build container). The reference itself has no tests, golden
vectors or fixtures, so these files -- outputs of the reference's own object code on the scenes of
tests/scenes.py -- are what pins the oracle on machines where oracle/_ref is absent.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refdrv  # noqa: E402
import scenes  # noqa: E402

# scene -> (builder kwargs, frames at which the state is recorded)
TRAJ = {
    "stack": ({}, [1, 10, 60, 240]),
    "brick_wall": ({}, [1, 30]),
    "cube_storm": ({}, [1, 60]),
    "seesaw": ({}, [1, 90]),
    "cube_and_ramp": ({}, [1, 90]),
    "coin": ({}, [1, 60]),
    "spring": ({}, [1, 60]),
    "hinge_joints": ({}, [1, 60]),
    "arm": ({}, [1, 60]),
    "triple_pendula": ({}, [1, 20]),
    "mirror_cube": ({}, [1, 90]),
    "spheres": ({}, [1, 120]),
    "pile": (dict(n_side=3), [1, 40]),
    "tumble": ({}, [1, 90]),
    "w256": ({}, [1, 5]),
    "spot_storm": (dict(n=2), [1, 24]),
}
HULLS = ["cube", "floor", "ico", "ramp", "cylinder", "lever", "seesaw_support"]


def main():
    out = {}
    for name, (kw, frames) in TRAJ.items():
        sc = scenes.BUILDERS[name](**kw)
        w = refdrv.RefWorld("strict").load(sc)
        w.log_enable(True)
        done = 0
        for f in frames:
            while done < f:
                w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
                done += 1
            out["%s/state/%d" % (name, f)] = w.state()
            out["%s/prev_vel/%d" % (name, f)] = w.prev_velocities()
            if f == 1:
                calls, contacts = w.log_get()
                out["%s/calls" % name] = calls
                out["%s/contacts" % name] = contacts
                w.log_enable(False)
        out["%s/params" % name] = w.params()
    np.savez_compressed(os.path.join(HERE, "trajectories.npz"), **out)

    hulls = {}
    for m in HULLS:
        sc = scenes.Scene("h")
        sc.bodies.append(scenes.BodyDesc((0, 0, 0), scenes.IDENT, 1.0, False, [scenes.hull(m, (1.0, 1.0, 1.0))]))
        w = refdrv.RefWorld("strict").load(sc)
        for k, v in w.hull(0).items():
            hulls["%s/%s" % (m, k)] = v
    np.savez_compressed(os.path.join(HERE, "hulls.npz"), **hulls)
    print("wrote", len(out), "trajectory arrays and", len(hulls), "hull arrays")


if __name__ == "__main__":
    main()
