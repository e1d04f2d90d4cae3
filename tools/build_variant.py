#!/usr/bin/env python
"""Experiment builds: exp/<name>/libpbg.so = the current sources with extra nvcc flags for the named translation units
(the other units are taken from the main build's object cache).

  python tools/build_variant.py phase ant,humanoid -DPBG_PHASE_CLOCKS
  python tools/build_variant.py old_ant ant --header /tmp/old/pbg_kernels.cuh     (another kernel header for those units)
"""
import os, shutil, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybullet_gym_b200 import _lib


def main():
    name, units = sys.argv[1], sys.argv[2].split(",")
    flags = [a for a in sys.argv[3:] if a.startswith("-") and a != "--header"]
    header = sys.argv[sys.argv.index("--header") + 1] if "--header" in sys.argv else None
    _lib.build_extension()
    out = os.path.join(ROOT, "exp", name)
    os.makedirs(out, exist_ok=True)
    src_dir = _lib.CSRC
    if header:
        src_dir = tempfile.mkdtemp()
        for f in os.listdir(_lib.CSRC):
            if f.endswith((".cu", ".cuh")):
                shutil.copy(os.path.join(_lib.CSRC, f), src_dir)
        shutil.copy(header, os.path.join(src_dir, "pbg_kernels.cuh"))
        os.makedirs(os.path.join(src_dir, "..", "..", "include"), exist_ok=True)
    objs = []
    jobs = []
    for f in sorted(os.listdir(_lib.CSRC)):
        if not f.endswith(".cu"):
            continue
        base = f[:-3]
        if base == "pbg_abi" or base[len("pbg_k_"):] in units:
            o = os.path.join(out, base + ".o")
            cmd = ["nvcc"] + _lib.NVCC_FLAGS + flags + ["-I", os.path.join(ROOT, "pybullet_gym_b200", "csrc"), "-c", "-o", o,
                                                        os.path.join(src_dir if base != "pbg_abi" else _lib.CSRC, f)]
            jobs.append(cmd)
            objs.append(o)
        else:
            objs.append(os.path.join(_lib.BUILD_DIR, base + ".o"))
    with ThreadPoolExecutor(8) as ex:
        for r in ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
            if r.returncode:
                sys.exit(r.stdout + r.stderr)
            if "-Xptxas" in flags:
                sys.stderr.write(r.stderr)
    lib = os.path.join(out, "libpbg.so")
    subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    print(lib)


if __name__ == "__main__":
    main()
