"""In-tree build of libmantaprover.so (sm_100a only) and of the CPU oracle used by the tests.

    python manta-rs_b200/build.py [--force] [--oracle]

Every .cu under csrc/ is compiled by its own nvcc process (in parallel), then linked with the static CUDA
runtime into manta-rs_b200/libmantaprover.so.  Objects are rebuilt when the source or any header is newer.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmantaprover.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _newest_header():
    t = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h", ".inc")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def _compile(src, obj, verbose):
    cmd = [NVCC, *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build_lib(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _newest_header()
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs, objs = [], []
    for f in srcs:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, out in ex.map(lambda j: _compile(j[0], j[1], verbose), jobs):
                if verbose or rc != 0:
                    sys.stderr.write(out)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


def build_oracle(force=False):
    """Compile the C++ restatement (test infrastructure) with its own Makefile."""
    odir = os.path.join(ROOT, "oracle")
    if os.path.exists(os.path.join(odir, "Makefile")):
        subprocess.run(["make", "-C", odir] + (["-B"] if force else []), check=True, capture_output=True)


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_lib(force=force, verbose="-v" in sys.argv))
    if "--oracle" in sys.argv:
        build_oracle(force)
