"""Per-kernel device time of any tests/scenes.py scene as a batch: `python scripts/scene_profile.py brick_wall 1 30 rows=32 cols=32 [coloured]`
(scene, worlds, frames, builder kwargs). Prints ms/frame and the per-kernel-family split of 10 further frames."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import __graft_entry__ as ge
import scenes

pkg = ge.load_package()
name, worlds, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[4:] if "=" in a)}
coloured = "coloured" in sys.argv[4:]
sc = scenes.BUILDERS[name](**kw)
b = pkg.Batch(pkg.Scene(sc), n_worlds=worlds, device=0, coloured=coloured)
b.set_scene_forces(sc)
if sc.initial_state is not None:
    b.broadcast(pkg.state15_to_21(sc.initial_state))
b.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
b.sync()
t0 = time.perf_counter()
for _ in range(frames):
    b.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
b.sync()
dt = time.perf_counter() - t0
c = b.counters()
print("%s x %d worlds (%d bodies), %s order: %.3f ms/frame, sweep depth %.0f, status 0x%x" % (
    name, worlds, len(sc.bodies), "coloured" if coloured else "reference", 1e3 * dt / frames, c["levels"] / c["frames"] / worlds,
    int(np.bitwise_or.reduce(b.status()))))
fam = b.profile(10, substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
print("   per-kernel ms over 10 more frames:", {k: round(v, 2) for k, v in fam.items()})
