"""Worker of tests/test_exactq.py: steps scenes with whatever library RAWPHYS_B200_LIB names (the RP_EXACT_QUATERNIONS build) in a
process of its own -- the package binds one library per process -- and saves the recorded states.
    python tests/exactq_worker.py OUT.npz scene:frame,frame ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import scenes  # noqa: E402

pkg = ge.load_package()
assert pkg.LIB_PATH.endswith("_exactq.so"), pkg.LIB_PATH
out = {}
for spec in sys.argv[2:]:
    name, frames = spec.split(":")
    frames = [int(f) for f in frames.split(",")]
    sc = scenes.BUILDERS[name]()
    b = pkg.Batch(pkg.Scene(sc), n_worlds=3, device=0)
    b.set_scene_forces(sc)
    for f in range(1, max(frames) + 1):
        b.step(1.0 / 60.0, sc.substeps, sc.iters, sc.collisions)
        if f in frames:
            st = b.state()
            assert np.array_equal(st[0], st[2])
            out["%s/state/%d" % (name, f)] = st[1][:, :15]
    out["%s/status" % name] = b.status()
    b.close()
np.savez(sys.argv[1], **out)
