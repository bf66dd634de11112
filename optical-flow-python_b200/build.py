#!/usr/bin/env python
"""Build libb200flow.so in-tree with nvcc for sm_100a (B200).  Usage: python build.py [--force] [--verbose]

Cross-compiles without a GPU.  -fmad=false: the reference is NumPy (no fused multiply-add); the integer-valued
decisions on the path (gray quantisation, pyramid sample indices, out-of-bounds tests) must round identically.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200flow.so")
SOURCES = ["pre.cu", "warp.cu", "solve.cu", "solve_ic.cu", "filter.cu", "eval.cu", "pipeline.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "b200flow.h"))
    headers.append(os.path.abspath(__file__))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((s, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r.returncode, r.stdout + r.stderr

    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=7) as ex:
        for name, rc, out in ex.map(run, jobs):
            if rc != 0 or verbose:
                sys.stderr.write("== %s (rc=%d)\n%s\n" % (name, rc, out))
            failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
