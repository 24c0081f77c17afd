"""Builds raw-physics_b200/librawphys_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

FP contract: --fmad=false (no FMA contraction), IEEE division and square root, no flush-to-zero -- the reference's
trajectories only reproduce under its own rounding (SURVEY.md TL;DR 3).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librawphys_b200.so")
SOURCES = [os.path.join(CSRC, "rp_batch.cu"), os.path.join(CSRC, "rp_scene.cpp"), os.path.join(CSRC, "rp_examples.cpp")]
HEADERS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))] + [
    os.path.join(os.path.dirname(HERE), "include", "rawphys_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-Xptxas", "-v",
]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False, out=None, defines=()):
    """`out`/`defines` build an experimental variant next to the product library (tuning runs only)."""
    if out is None and not force and up_to_date():
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out or OUT] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    log_name = "build.log" if out is None else "build_%s.log" % os.path.splitext(os.path.basename(out))[0]  # (variants build side by side)
    with open(os.path.join(HERE, log_name), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed (see raw-physics_b200/%s)" % log_name)
    if verbose:
        print(log)
    return out or OUT


# The reference's compile-time switch USE_QUATERNIONS_LINEARIZED_FORMULAS (pbd.cpp:16, pbd_base_constraints.cpp:4) is a compile-time
# switch here too: -DRP_EXACT_QUATERNIONS builds the non-linearised orientation updates (axis-angle through libm sin / cos) into a
# second library with the same C ABI. Load it with RAWPHYS_B200_LIB=.../librawphys_b200_exactq.so or link against it.
EXACTQ_OUT = os.path.join(HERE, "librawphys_b200_exactq.so")


def build_exactq(force=False):
    if not force and os.path.exists(EXACTQ_OUT) and all(os.path.getmtime(f) <= os.path.getmtime(EXACTQ_OUT) for f in SOURCES + HEADERS + [os.path.abspath(__file__)]):
        return EXACTQ_OUT
    return build(out=EXACTQ_OUT, defines=("RP_EXACT_QUATERNIONS",))


# Single precision (SURVEY.md 7 hard part 1, 8d "fast mode"): the same sources with `real` = float (rp_math.h). Same C ABI (host
# records stay double, converted at upload / download). An experiment, not a product mode: see DESIGN.md 7 for what holds and what does not.
F32_OUT = os.path.join(HERE, "librawphys_b200_f32.so")


def build_f32(force=False):
    if not force and os.path.exists(F32_OUT) and all(os.path.getmtime(f) <= os.path.getmtime(F32_OUT) for f in SOURCES + HEADERS + [os.path.abspath(__file__)]):
        return F32_OUT
    return build(out=F32_OUT, defines=("RP_REAL_F32",))


def build_all(force=False):
    """the product library and its two build variants, compiled side by side (three nvcc processes)"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(3) as ex:
        jobs = [ex.submit(build, force), ex.submit(build_exactq, force), ex.submit(build_f32, force)]
        return [j.result() for j in jobs]


HOST_SRC = os.path.join(HERE, "host", "rp_headless.cpp")
HOST_OUT = os.path.join(HERE, "rp_headless")


def build_host(force=False):
    """The headless driver (host C++ over the C ABI): raw-physics_b200/rp_headless, linked against the library next to it."""
    deps = [HOST_SRC, OUT, os.path.join(CSRC, "rp_math.h"), os.path.join(os.path.dirname(HERE), "include", "rawphys_b200.h")]
    if not force and os.path.exists(HOST_OUT) and all(os.path.getmtime(f) <= os.path.getmtime(HOST_OUT) for f in deps):
        return HOST_OUT
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-o", HOST_OUT, HOST_SRC, "-L" + HERE, "-lrawphys_b200",
           "-Wl,-rpath,$ORIGIN", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building rp_headless failed")
    return HOST_OUT


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    build_host(force="--force" in sys.argv)
