"""Config 4 in small: a pile of random hulls (ico / cylinder / cube / sphere, tests/scenes.py::pile) as ONE scene, stepped in
both solve orders; prints ms/frame and the work counters. `python scripts/pile_run.py [n_side] [frames]` on a GPU box."""
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
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 10
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sc = scenes.pile(n_side=n_side)
for coloured in (False, True):
    b = pkg.Batch(pkg.Scene(sc), n_worlds=1, device=0, coloured=coloured)
    b.set_scene_forces(sc)
    b.step()
    b.sync()
    t0 = time.perf_counter()
    for _ in range(frames):
        b.step()
    b.sync()
    dt = time.perf_counter() - t0
    c = b.counters()
    st = b.state()[0]
    print("pile %d^3 = %d bodies, %s order: %.2f ms/frame, sweep depth %.0f, contacts/substep %.0f, status 0x%x, lowest y %.3f, max speed %.2f" % (
        n_side, len(sc.bodies), "coloured" if coloured else "reference", 1e3 * dt / frames, c["levels"] / c["frames"],
        c["contacts"] / (c["frames"] * 20.0), int(np.bitwise_or.reduce(b.status())), st[1:, 1].min(), np.sqrt((st[1:, 7:10] ** 2).sum(1)).max()))
    fam = b.profile(10)
    print("   per-kernel ms over 10 more frames:", {k: round(v, 2) for k, v in fam.items()})
    b.close()
