"""Offline viewer / exporter: turns state dumps of the headless driver into pictures.

The reference's only "test" is looking at its GLFW window (README GIFs, render/graphics.cpp:127-151 draws every entity's
mesh with model matrix T * R * S, entity.cpp:123-145). There is no window here; instead

    rp_headless --scene stack --frames 120 --dump run.rphd [--dump-worlds 4]
    python -m viewer --scene stack --dump run.rphd --out run.gif          (from raw-physics_b200/; or scripts/render_dump.py)

draws the collider hulls (their face loops, i.e. exactly the geometry the solver sees) of every body at every dumped frame
with a fixed perspective camera into an animated GIF, one tile per dumped world, plus an optional .obj export of single
frames for external tools. Geometry comes from the library's own scene builders (rp_example_create + rp_scene_hull_dump, host
code: no GPU needed to render). Pure numpy + PIL; nothing here is on the simulation path.
"""
import argparse
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_dump(path):
    """rp_headless --dump file -> (frames [R], states [R][worlds][bodies][stride]). Header: "RPHD", u32 version, u32 bodies,
    u32 records, u32 stride, u32 every, u32 worlds (version >= 2, else 1), u32 reserved."""
    with open(path, "rb") as f:
        head = f.read(32)
        magic, version, bodies, records, stride, every, worlds, _ = struct.unpack("<4sIIIIIII", head)
        if magic != b"RPHD" or version not in (1, 2):
            raise ValueError("%s: not an rp_headless dump" % path)
        if version == 1:
            worlds = 1
        frames = np.zeros(records, dtype=np.int64)
        states = np.zeros((records, worlds, bodies, stride))
        for r in range(records):
            tag = struct.unpack("<II", f.read(8))
            frames[r] = tag[0]
            states[r] = np.frombuffer(f.read(8 * worlds * bodies * stride), dtype="<f8").reshape(worlds, bodies, stride)
    return frames, states


def save_dump(path, frames, states, every=1):
    """the same format from Python (tests, Batch.state() recordings): states [R][worlds][bodies][stride]"""
    states = np.ascontiguousarray(states, dtype="<f8")
    R, W, B, S = states.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<4sIIIIIII", b"RPHD", 2 if W > 1 else 1, B, R, S, every, W, 0))
        for r in range(R):
            f.write(struct.pack("<II", int(frames[r]), 0))
            f.write(states[r].tobytes())


def scene_geometry(scene):
    """per body: list of colliders, each ("hull", vertices [V][3], face loops [list of index arrays]) or ("sphere", radius)"""
    import ctypes as C
    out = []
    for b in range(scene.n):
        o = np.zeros(16)
        scene.L.rp_scene_body_desc(scene.h, b, o.ctypes.data_as(C.POINTER(C.c_double)))
        cols = []
        for c in range(int(o[12])):
            h = scene.hull(b, c)
            if h is None:
                nv, ni, rad = C.c_uint32(), C.c_uint32(), C.c_float()
                scene.L.rp_scene_collider_soup_size(scene.h, b, c, C.byref(nv), C.byref(ni), C.byref(rad))
                cols.append(("sphere", float(rad.value)))
            else:
                loops = [h["face_idx"][h["face_ptr"][i]:h["face_ptr"][i + 1]].astype(np.int64) for i in range(len(h["face_ptr"]) - 1)]
                cols.append(("hull", h["verts"].copy(), loops))
        out.append(cols)
    return out


def quat_rotate(q, v):
    """rows of v rotated by the unit quaternion q = (x, y, z, w)"""
    u = q[:3]
    return v + 2.0 * np.cross(u, np.cross(u, v) + q[3] * v)


def _sphere_rings(radius, n=20):
    t = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    c, s, z = np.cos(t) * radius, np.sin(t) * radius, np.zeros(n)
    return [np.stack([c, s, z], 1), np.stack([c, z, s], 1), np.stack([z, c, s], 1)]


def world_polylines(geometry, state):
    """closed polylines (arrays [n][3]) of every collider of one world's state [bodies][stride]"""
    lines = []
    for b, cols in enumerate(geometry):
        x, q = state[b, 0:3], state[b, 3:7]
        for col in cols:
            if col[0] == "hull":
                wv = quat_rotate(q, col[1]) + x
                lines.extend(wv[loop] for loop in col[2])
            else:
                lines.extend(quat_rotate(q, ring) + x for ring in _sphere_rings(col[1]))
    return lines


class Camera:
    """pinhole camera looking at `target` from `eye` (y up), like the reference's default view of its examples"""

    def __init__(self, eye, target, fov_deg=45.0):
        eye, target = np.asarray(eye, float), np.asarray(target, float)
        f = target - eye
        f /= np.linalg.norm(f)
        r = np.cross(f, [0.0, 1.0, 0.0])
        r /= np.linalg.norm(r)
        self.eye, self.f, self.r, self.u = eye, f, r, np.cross(r, f)
        self.k = 1.0 / np.tan(np.radians(fov_deg) / 2.0)

    def project(self, p, size):
        d = p - self.eye
        z = np.maximum(d @ self.f, 1e-3)
        x = (d @ self.r) / z * self.k
        y = (d @ self.u) / z * self.k
        return np.stack([(x + 1.0) * 0.5 * size, (1.0 - y) * 0.5 * size], 1), d @ self.f


def auto_camera(geometry, states):
    """frames the bounding box of every dynamic body's positions over the whole recording"""
    pos = states[..., 0:3].reshape(-1, 3)
    pos = pos[np.isfinite(pos).all(1)]
    lo, hi = np.percentile(pos, 1, axis=0), np.percentile(pos, 99, axis=0)
    centre = (lo + hi) / 2.0
    extent = max(float(np.max(hi - lo)), 4.0)
    return Camera(centre + np.array([0.6, 0.45, 1.0]) * extent * 1.4, centre)


def render_gif(geometry, frames, states, path, size=320, camera=None, max_worlds=4, duration_ms=50):
    """animated GIF, one tile per world (first max_worlds), one image per record; returns the number of images"""
    from PIL import Image, ImageDraw
    R, W = states.shape[0], min(states.shape[1], max_worlds)
    cam = camera or auto_camera(geometry, states[:, :W])
    images = []
    for r in range(R):
        img = Image.new("RGB", (size * W, size), (18, 18, 24))
        for w in range(W):
            tile = Image.new("RGB", (size, size), (18, 18, 24))  # (a tile of its own: long edges are clipped at its border)
            draw = ImageDraw.Draw(tile)
            for line in world_polylines(geometry, states[r, w]):
                xy, depth = cam.project(line, size)
                if (depth < 0.05).any() or not np.isfinite(xy).all():
                    continue
                pts = [(float(a), float(b)) for a, b in xy]
                draw.line(pts + pts[:1], fill=(120, 200, 255), width=1)
            img.paste(tile, (w * size, 0))
        ImageDraw.Draw(img).text((4, 4), "frame %d" % int(frames[r]), fill=(255, 255, 255))
        images.append(img)
    images[0].save(path, save_all=True, append_images=images[1:], duration=duration_ms, loop=0)
    return len(images)


def export_obj(geometry, state, path):
    """one world's state as a Wavefront .obj of the collider hulls (polygon faces; spheres as three rings of line elements)"""
    nv = 0
    with open(path, "w") as f:
        for b, cols in enumerate(geometry):
            x, q = state[b, 0:3], state[b, 3:7]
            for ci, col in enumerate(cols):
                f.write("o body%d_collider%d\n" % (b, ci))
                if col[0] == "hull":
                    for p in quat_rotate(q, col[1]) + x:
                        f.write("v %.17g %.17g %.17g\n" % tuple(p))
                    for loop in col[2]:
                        f.write("f " + " ".join(str(nv + 1 + int(i)) for i in loop) + "\n")
                    nv += col[1].shape[0]
                else:
                    for ring in _sphere_rings(col[1]):
                        for p in quat_rotate(q, ring) + x:
                            f.write("v %.17g %.17g %.17g\n" % tuple(p))
                        f.write("l " + " ".join(str(nv + 1 + i) for i in list(range(len(ring))) + [0]) + "\n")
                        nv += len(ring)
    return nv


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--scene", required=True, help="built-in scene name (rp_example_create), as given to rp_headless")
    ap.add_argument("--params", default="", help="comma-separated scene parameters, as given to rp_headless")
    ap.add_argument("--perturb", action="store_true")
    ap.add_argument("--dump", required=True)
    ap.add_argument("--out", default="", help=".gif to write")
    ap.add_argument("--obj", default="", help=".obj to write (world 0 of the last record)")
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--worlds", type=int, default=4)
    a = ap.parse_args(argv)
    sys.path.insert(0, os.path.dirname(HERE))
    import __graft_entry__ as ge
    pkg = ge.load_package()
    name, perturb = (("hinge_joints", True) if a.scene == "levers" else (a.scene, a.perturb))
    params = [float(x) for x in a.params.split(",") if x]
    scene, _ = pkg.example(name, params, perturb=perturb)
    geometry = scene_geometry(scene)
    frames, states = load_dump(a.dump)
    if states.shape[2] != len(geometry):
        raise SystemExit("dump has %d bodies, scene %s has %d" % (states.shape[2], a.scene, len(geometry)))
    if a.out:
        n = render_gif(geometry, frames, states, a.out, size=a.size, max_worlds=a.worlds)
        print("%s: %d images, %d world(s)" % (a.out, n, min(states.shape[1], a.worlds)))
    if a.obj:
        print("%s: %d vertices" % (a.obj, export_obj(geometry, states[-1, 0], a.obj)))


if __name__ == "__main__":
    main()
