"""A reduced run of every kernel family for compute-sanitizer (scripts/gpu_sanitize.sh): few worlds, few frames, all code paths --
batched small worlds (level-major cooperative sweeps), world-block sweeps, joints, compound bodies + large hulls (warp-per-pair
kernels, warp clipping), one large scene (grid broadphase, union-find islands, parallel colouring), host-buffer steps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()


def run(name, params=(), perturb=False, worlds=3, frames=3, **kw):
    scene, desc = pkg.example(name, params, perturb=perturb)
    b = pkg.Batch(scene, n_worlds=worlds, device=0, **kw)
    b.set_scene_forces(desc)
    for _ in range(frames):
        b.step(1.0 / 60.0, desc.substeps, desc.iters, desc.collisions)
    st = b.state()
    buf = np.zeros_like(st)
    b.step_host(st.ctypes.data, buf.ctypes.data, 1.0 / 60.0, desc.substeps, desc.iters, desc.collisions)
    assert np.isfinite(buf).all() and not b.status().any(), (name, b.status())
    b.close()
    print("ok", name, kw, flush=True)


run("stack", worlds=5, frames=40)               # contacts form around frame 30 (fewer than 64 worlds: barrier sweeps)
run("stack", worlds=70, frames=36)              # 64 worlds and more: dataflow sweeps (pos_flow / vel_flow)
os.environ["RP_FLOW"] = "2"                     # ... forced on for a joint scene with contacts and for one large scene
run("seesaw", worlds=40, frames=30)
run("brick_wall", (6, 6), worlds=1, frames=30, coloured=True)
del os.environ["RP_FLOW"]
scene, desc = pkg.example("spot_storm", (1, 7), hull_device=0)   # device-side hull construction (rp_hull.cuh), 529-vertex hull included
assert scene.hull_build_stats()[0] >= 11
print("ok device hulls", scene.hull_build_stats(), flush=True)
run("w256", worlds=64, frames=2)
run("w256", (2, 2, 4), worlds=33, frames=45, sweep_block_worlds=4)
run("hinge_joints", perturb=True, worlds=40, frames=5)
run("hinge_joints", perturb=True, worlds=40, frames=5, sweep_block_worlds=8)
run("coin", worlds=2, frames=50)                 # 64-gon caps: warp kernels, warp clipping
run("spot_storm", (2, 99), worlds=2, frames=14, max_pairs=8192, max_contacts=8192)
run("brick_wall", (8, 8), worlds=1, frames=40, coloured=True)
run("pile", (6, 12345, 2.2), worlds=1, frames=40, coloured=True, large_scene=2)
run("pile", (5, 12345, 2.2), worlds=2, frames=40, large_scene=2)
print("sanitize driver done")
