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
    "w256": ({}, [1, 5, 20, 40, 60]),
    "spot_storm": (dict(n=2), [1, 24]),
    # round 2: the joint types / axis selectors no reference example steps, and the weld hinge at 50 x 50
    "rott_pendulum": ({}, [1, 30]),
    "mutual_orientation": ({}, [1, 60]),
    "negative_axes": ({}, [1, 60]),
}
# the timed window of bench.py's workloads: per-frame narrowphase call / contact counts over the whole window and a digest of
# the last frame's contact log (every call row and every contact point, in order), so that what is TIMED is pinned
WINDOWS = {"w256": ({}, 60), "brick_wall_32x32": (dict(rows=32, cols=32), 30)}
HULLS = ["cube", "floor", "ico", "ramp", "cylinder", "lever", "seesaw_support"]


def log_digest(calls, contacts):
    """sha256 over the call rows (a, b, count, first) and the contact points + normals of one frame's log, as raw bytes"""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(calls, dtype=np.uint32).tobytes())
    h.update(np.ascontiguousarray(contacts, dtype=np.float64).tobytes())
    return h.digest()


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

    win = {}
    for name, (kw, frames) in WINDOWS.items():
        sc = scenes.BUILDERS[name.split("_32")[0]](**kw)
        w = refdrv.RefWorld("strict").load(sc)
        w.log_enable(True)
        counts = np.zeros((frames, 2), dtype=np.int64)
        for f in range(frames):
            w.log_clear()
            w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
            calls, contacts = w.log_get()
            counts[f] = (len(calls), len(contacts))
            if f + 1 in (frames // 3, 2 * frames // 3, frames):
                win["%s/state/%d" % (name, f + 1)] = w.state()
        win["%s/counts" % name] = counts
        win["%s/last_log_digest" % name] = np.frombuffer(log_digest(calls, contacts), dtype=np.uint8)
        sub = refdrv.split_substeps(calls)
        first = int(sub[-1][0][3])  # contacts of the last substep of the last frame, kept in full
        win["%s/last_substep_calls" % name] = np.asarray(sub[-1], dtype=np.uint32)
        win["%s/last_substep_contacts" % name] = contacts[first:]
    np.savez_compressed(os.path.join(HERE, "windows.npz"), **win)

    hulls = {}
    for m in HULLS:
        sc = scenes.Scene("h")
        sc.bodies.append(scenes.BodyDesc((0, 0, 0), scenes.IDENT, 1.0, False, [scenes.hull(m, (1.0, 1.0, 1.0))]))
        w = refdrv.RefWorld("strict").load(sc)
        for k, v in w.hull(0).items():
            hulls["%s/%s" % (m, k)] = v
    np.savez_compressed(os.path.join(HERE, "hulls.npz"), **hulls)
    print("wrote", len(out), "trajectory arrays,", len(win), "window arrays and", len(hulls), "hull arrays")


if __name__ == "__main__":
    main()
