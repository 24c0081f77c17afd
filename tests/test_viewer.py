"""The offline viewer / exporter (raw-physics_b200/viewer.py): dump format round trip, geometry taken from the
library's scene builders, GIF and .obj output. Host-only: the states drawn here are the compiled reference's committed ones."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))


@pytest.fixture(scope="module")
def viewer(pkg):
    import importlib
    return importlib.import_module("rawphys_b200.viewer")


def recorded(name, frames, worlds=1, stride=21):
    st = []
    for f in frames:
        g = GOLD["%s/state/%d" % (name, f)]
        rec = np.zeros((worlds, g.shape[0], stride))
        rec[:, :, :g.shape[1]] = g
        st.append(rec)
    return np.array(frames), np.stack(st)


@pytest.mark.parametrize("worlds", [1, 3])
def test_dump_round_trip(viewer, tmp_path, worlds):
    frames, states = recorded("stack", [1, 10, 60], worlds)
    path = str(tmp_path / "d.rphd")
    viewer.save_dump(path, frames, states)
    f2, s2 = viewer.load_dump(path)
    assert np.array_equal(f2, frames) and np.array_equal(s2, states)
    with pytest.raises(ValueError):
        open(path, "r+b").write(b"XXXX")
        viewer.load_dump(path)


def test_geometry_is_the_solvers(pkg, viewer):
    scene, _ = pkg.example("stack")
    geo = viewer.scene_geometry(scene)
    assert len(geo) == 9 and all(len(c) == 1 and c[0][0] == "hull" for c in geo)
    kind, verts, loops = geo[1][0]
    assert verts.shape == (8, 3) and len(loops) == 6 and all(len(l) == 4 for l in loops)
    scene, _ = pkg.example("spheres") if "spheres" in pkg.example_names() else (None, None)
    if scene is not None:
        assert any(c[0][0] == "sphere" and c[0][1] > 0 for c in viewer.scene_geometry(scene))


def test_gif_and_obj(pkg, viewer, tmp_path):
    from PIL import Image
    scene, _ = pkg.example("stack")
    geo = viewer.scene_geometry(scene)
    frames, states = recorded("stack", [1, 10, 60, 240], worlds=2)
    gif = str(tmp_path / "stack.gif")
    assert viewer.render_gif(geo, frames, states, gif, size=160) == 4
    im = Image.open(gif)
    assert im.n_frames == 4 and im.size == (320, 160)
    im.seek(3)
    px = np.asarray(im.convert("RGB"))
    assert (px.max(axis=2) > 100).sum() > 200          # the wireframes are there
    assert (px[24:, :160] == px[24:, 160:]).mean() > 0.999  # identical worlds draw identical tiles (below the frame label)
    obj = str(tmp_path / "stack.obj")
    assert viewer.export_obj(geo, states[-1, 0], obj) == 9 * 8
    text = open(obj).read()
    assert text.count("\nf ") == 9 * 6 and text.count("o body") == 9
