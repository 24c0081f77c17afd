"""Generates tests/golden/exactq.npz from the reference compiled WITHOUT its USE_QUATERNIONS_LINEARIZED_FORMULAS switch
(oracle/Makefile `ref_exactq`: the two #define lines filtered out of build-time copies, everything else unmodified; needs
/root/reference, i.e. the build container). Pins the RP_EXACT_QUATERNIONS builds of the restatement and of the product.

    python tests/golden/make_golden_exactq.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refdrv  # noqa: E402
import scenes  # noqa: E402

TRAJ = {"stack": [1, 10, 40, 60], "seesaw": [1, 60], "cube_and_ramp": [1, 60], "tumble": [1, 30, 60], "spring": [1, 60]}


def main():
    out = {}
    for name, frames in TRAJ.items():
        sc = scenes.BUILDERS[name]()
        w = refdrv.RefWorld("strict_exactq").load(sc)
        for f in range(1, max(frames) + 1):
            w.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
            if f in frames:
                out["%s/state/%d" % (name, f)] = w.state()
        print(name, "done")
    np.savez_compressed(os.path.join(HERE, "exactq.npz"), **out)


if __name__ == "__main__":
    main()
