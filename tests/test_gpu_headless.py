"""The C++ headless driver (raw-physics_b200/host/rp_headless.cpp -> raw-physics_b200/rp_headless): scenes built and
stepped by host C++ through the C ABI only, its state dumps compared with the oracle stepping the same example
(tests/scenes.py restates the same init() halves independently, in Python)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import refdrv
import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "raw-physics_b200", "rp_headless")


def read_dump(path):
    """-> {frame: [bodies][21] array} (format: header of rp_headless.cpp)."""
    raw = open(path, "rb").read()
    magic, version, nb, records, stride, every, _ = struct.unpack_from("<IIIIIIQ", raw, 0)
    assert magic == 0x44485052 and version == 1 and stride == 21
    out, off = {}, 32
    for _ in range(records):
        frame, _pad = struct.unpack_from("<II", raw, off)
        off += 8
        out[frame] = np.frombuffer(raw, dtype="<f8", count=nb * stride, offset=off).reshape(nb, stride).copy()
        off += nb * stride * 8
    assert off == len(raw)
    return out


def run(tmp_path, *args):
    dump = str(tmp_path / "dump.bin")
    r = subprocess.run([EXE, "--dump", dump] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1]), read_dump(dump)


def test_headless_driver_needs_a_gpu(pkg):
    """No CUDA device -> a clear failure, never a CPU path (runs on the CPU-only build container too)."""
    assert os.path.exists(EXE)
    r = subprocess.run([EXE, "--scene", "stack", "--frames", "1"], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("scene,args,builder,kw,frames", [
    ("stack", [], "stack", {}, 90),
    ("brick_wall", ["--rows", 6, "--cols", 4], "brick_wall", dict(rows=6, cols=4), 60),
    ("w256", [], "w256", {}, 12),
])
def test_headless_dump_bit_exact(pkg, oracle_flavour, tmp_path, scene, args, builder, kw, frames):
    info, dump = run(tmp_path, "--scene", scene, "--frames", frames, "--worlds", 5, "--dump-every", 3, *args)
    assert info["status_bits"] == 0 and info["diverged_worlds"] == 0 and info["worlds"] == 5
    sc = scenes.BUILDERS[builder](**kw)
    assert info["bodies"] == len(sc.bodies)
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    seen = 0
    for f in range(1, frames + 1):
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        if f in dump:
            assert np.array_equal(dump[f][:, :15], o.state()), (scene, f)
            seen += 1
    assert seen == len(dump) == frames // 3


@pytest.mark.gpu
def test_headless_levers_within_tolerance(pkg, oracle_flavour, tmp_path):
    """hinge_joints.cpp's levers, spun about their hinges: libm (asin/sin/cos) scene, 1e-9 as in test_gpu_parity.py."""
    info, dump = run(tmp_path, "--scene", "levers", "--frames", 60, "--worlds", 2)
    assert info["status_bits"] == 0 and info["diverged_worlds"] == 0
    sc = scenes.BUILDERS["hinge_joints"]()
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    worst = 0.0
    for f in range(1, 61):
        o.step(substeps=sc.substeps, iters=sc.iters, collisions=sc.collisions)
        worst = max(worst, float(np.abs(dump[f][:, :7] - o.state()[:, :7]).max()))
    print("levers: worst |pose diff| = %g" % worst)
    assert worst <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("scene,extra,frames,exact", [
    ("cube_storm", [], 60, True), ("coin", [], 60, True), ("seesaw", [], 60, True), ("mirror_cube", [], 60, True), ("spring", [], 60, True),
    ("cube_and_ramp", [], 60, True), ("spot_storm", ["--params", "2,99"], 12, True), ("debug", [], 5, True),
    ("arm", ["--perturb"], 40, False), ("rott_pendulum", ["--perturb"], 20, False), ("triple_pendula", ["--perturb"], 20, False),
])
def test_headless_runs_every_example(pkg, oracle_flavour, tmp_path, scene, extra, frames, exact):
    """the init()/update() halves of the remaining examples in host C++ (rp_example_create + the frame loop of
    rp_headless), each with the step settings its own update() uses, against the oracle stepping the same description."""
    info, dump = run(tmp_path, "--scene", scene, "--frames", frames, "--worlds", 3, "--dump-every", frames, *extra)
    assert info["status_bits"] == 0 and info["diverged_worlds"] == 0
    params = [float(x) for x in extra[extra.index("--params") + 1].split(",")] if "--params" in extra else ()
    _, desc = pkg.example(scene, params, perturb="--perturb" in extra)
    o = refdrv.RefWorld(oracle_flavour).load(desc)
    for _ in range(frames):
        o.step(substeps=desc.substeps, iters=desc.iters, collisions=desc.collisions)
    got, want = dump[frames][:, :15], o.state()
    if exact:
        assert np.array_equal(got, want), (scene, np.abs(got - want).max())
    else:
        assert np.abs(got[:, :7] - want[:, :7]).max() <= 1e-9


@pytest.mark.gpu
def test_headless_shards_over_two_gpus(pkg, oracle_flavour, tmp_path):
    """--gpus 2: the worlds in two contiguous shards, one batch per device, every frame enqueued on both before either is
    waited for; the shards agree with each other bit for bit and with the oracle. Needs two devices."""
    if pkg.lib().rp_device_count() < 2:
        pytest.skip("one CUDA device")
    info, dump = run(tmp_path, "--scene", "stack", "--frames", 45, "--worlds", 65, "--gpus", 2, "--dump-every", 45)
    assert info["status_bits"] == 0 and info["diverged_worlds"] == 0 and info["gpus"] == 2 and info["worlds"] == 65
    sc = scenes.stack()
    o = refdrv.RefWorld(oracle_flavour).load(sc)
    for _ in range(45):
        o.step()
    assert np.array_equal(dump[45][:, :15], o.state())


@pytest.mark.gpu
def test_headless_device_hulls_and_multi_world_dump(pkg, oracle_flavour, tmp_path):
    """--device-hulls (hull topology built by csrc/rp_hull.cuh) and --dump-worlds (8 f4: a version-2 dump of
    several worlds, read back and drawn by raw-physics_b200/viewer.py): coin.cpp's 64-gon cylinder caps, bit-exact against
    the oracle, every dumped world identical."""
    import importlib
    viewer = importlib.import_module("rawphys_b200.viewer")
    dump = str(tmp_path / "coin.rphd")
    r = subprocess.run([EXE, "--scene", "coin", "--frames", "40", "--worlds", "4", "--dump", dump, "--dump-every", "10", "--dump-worlds", "3",
                        "--device-hulls"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["hull_builder"] == "device" and info["hulls_built"] >= 2 and info["status_bits"] == 0
    frames, states = viewer.load_dump(dump)
    assert list(frames) == [10, 20, 30, 40] and states.shape[1] == 3
    assert np.array_equal(states[:, 0], states[:, 1]) and np.array_equal(states[:, 0], states[:, 2])
    scene, desc = pkg.example("coin")
    o = refdrv.RefWorld(oracle_flavour).load(desc)
    for _ in range(40):
        o.step(substeps=desc.substeps, iters=desc.iters, collisions=desc.collisions)
    assert np.array_equal(states[-1, 0][:, :15], o.state())
    gif = str(tmp_path / "coin.gif")
    assert viewer.render_gif(viewer.scene_geometry(scene), frames, states, gif, size=128) == 4
    assert os.path.getsize(gif) > 1000
