"""Build libscape_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "kernels.cu", "em_cluster.cu", "em_tail.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in ("kernels.cuh", "em_device.cuh", "host_prep.hpp", "np_rng.hpp", "work_pool.hpp")] + \
    [os.path.join(os.path.dirname(HERE), "include", "scape_b200.h")]
LIB = os.path.join(HERE, "libscape_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-O3,-pthread", "-shared", "-cudart", "static"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the translation units in parallel (objects under csrc/_obj/), then link."""
    if not force and not needs_build():
        return LIB
    obj_dir = os.path.join(HERE, "csrc", "_obj")
    os.makedirs(obj_dir, exist_ok=True)
    cflags = [f for f in FLAGS if f not in ("-shared",)] + (["-Xptxas", "-v"] if verbose else [])
    procs = []
    for src in SRC:
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        deps = [src] + [d for d in DEPS if not d.endswith(".cu")]
        fresh = (not force and os.path.exists(obj) and
                 all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps))
        cmd = [NVCC] + cflags + ["-c", "-o", obj, src]
        procs.append((obj, cmd, None if fresh else subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                                    stderr=subprocess.STDOUT, text=True)))
    for obj, cmd, pr in procs:
        if pr is None:
            continue
        out, _ = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        if verbose:
            sys.stderr.write(out)
    tmp = LIB + ".tmp%d" % os.getpid()      # link beside the target, then rename: the .so is never half-written
    cmd = [NVCC] + FLAGS + ["-o", tmp] + [o for o, _, _ in procs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc link failed: " + " ".join(cmd))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
