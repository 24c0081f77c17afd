"""Hull construction, host (build_hull, rp_scene.cpp) against device (rp_hull.cuh): milliseconds per mesh, as
rp_scene_hull_build_stats reports them (host: wall clock of the build; device: CUDA events around the kernels, copies of the
result excluded). `python scripts/hull_build_time.py` on a GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import scenes

pkg = ge.load_package()
meshes = sorted(f[:-4] for f in os.listdir(os.path.join(ROOT, "tests", "golden", "meshes")) if f.endswith(".f32"))


def one(mesh, dev):
    sc = scenes.Scene("h")
    sc.bodies.append(scenes.BodyDesc((0, 0, 0), scenes.IDENT, 1.0, False, [scenes.hull(mesh, (1.0, 1.0, 1.0))]))
    s = pkg.Scene(sc, hull_device=dev)
    h = s.hull(0)
    return s.hull_build_stats()[1], h["verts"].shape[0], h["normals"].shape[0]


one("cube", 0)  # context + module load
out = {}
for m in meshes:
    host = min(one(m, -1)[0] for _ in range(3))
    dev, V, F = min(one(m, 0) for _ in range(3))
    out[m] = dict(vertices=V, faces=F, host_ms=round(host, 3), device_ms=round(dev, 3))
    print("%-24s V %4d F %4d   host %8.3f ms   device %8.3f ms" % (m, V, F, host, dev))
print(json.dumps(out))
