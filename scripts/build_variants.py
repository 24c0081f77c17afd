"""Builds tuning variants of the library next to the product: raw-physics_b200/variants/NAME.so with extra -D defines.
    python scripts/build_variants.py NAME=DEF1,DEF2=3 [NAME2=...]
Variants are scratch (git-ignored); they travel to the GPU box and are selected with RAWPHYS_B200_LIB."""
import importlib.util
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "raw-physics_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
os.makedirs(os.path.join(ROOT, "raw-physics_b200", "variants"), exist_ok=True)


def one(arg):
    name, _, defs = arg.partition("=")
    out = os.path.join(ROOT, "raw-physics_b200", "variants", name + ".so")
    import subprocess
    cmd = [b.nvcc()] + b.NVCC_FLAGS + ["-D" + d for d in defs.split(",") if d] + ["-o", out] + b.SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    open(out + ".log", "w").write(r.stdout + r.stderr)
    return name, r.returncode


with ThreadPoolExecutor(4) as ex:
    for name, rc in ex.map(one, sys.argv[1:]):
        print(name, "ok" if rc == 0 else "FAILED (see variants/%s.so.log)" % name)
