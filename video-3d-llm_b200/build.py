"""Build libv3d.so (the C ABI of include/v3d.h) in-tree with nvcc for sm_100a.

    python video-3d-llm_b200/build.py [--force] [--verbose]

The library travels to the GPU box with the repository snapshot (it is git-ignored, not
gpurun-ignored), so nothing is compiled there.  The CUDA runtime is linked statically: the
only load-time dependencies are libc/libstdc++ and the driver's libcuda.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libv3d.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def up_to_date() -> bool:
    if not os.path.isfile(LIB):
        return False
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(os.path.dirname(HERE), "include", "v3d.h")]
    return os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines/out: build an experimental variant (e.g. -DV3D_FUSED_MIN_BLOCKS=4) next to the default library."""
    global LIB
    if out:
        LIB = os.path.join(OUT_DIR, out)
        force = True
    if not force and up_to_date():
        return LIB
    nvcc = find_nvcc()
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(os.path.dirname(HERE), "include", "v3d.h"), os.path.abspath(__file__)]
    hdr_time = max(os.path.getmtime(h) for h in headers)
    for src in sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + (("." + out) if out else "") + ".o")
        objs.append(obj)
        # objects are reused when neither their source nor any header changed (default flags only)
        if not defines and not out and not verbose and os.environ.get("V3D_FORCE_BUILD", "0") != "1" and os.path.isfile(obj) \
                and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
            continue
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(out)
        if p.returncode != 0:
            failed = True
            print(f"nvcc failed on {src}", file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", LIB, *objs]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
