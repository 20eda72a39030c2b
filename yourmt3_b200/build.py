"""In-tree build of the C-ABI CUDA library for sm_100a (explicit nvcc, no JIT cache).

``python -m yourmt3_b200.build`` or ``__graft_entry__.build()``.  Objects go to
``build/`` (git-ignored); the shared library lands next to the sources so it
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "yourmt3_b200", "csrc")
OUT = os.environ.get("YMT3_B200_LIB") or os.path.join(CSRC, "libymt3_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj_debug" if os.environ.get("YMT3_B200_LIB") else "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include")]
CFLAGS += os.environ.get("YMT3_EXTRA_NVCC_FLAGS", "").split()   # debug builds only (e.g. -DYMT3_GEMM_TRACE)


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "ymt3_b200.h"))
    return sorted(hs)


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build_native(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_digest = _digest(_headers(), " ".join(ARCH + CFLAGS))
    objs, jobs = [], []
    for src in _sources():
        tag = _digest([src], hdr_digest)
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + "." + tag + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append([NVCC, *ARCH, *CFLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else []))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out.strip():
                    print(out)
    # prune stale objects of earlier source revisions (the directory travels with gpurun snapshots)
    keep = {os.path.basename(o) for o in objs}
    for f in os.listdir(OBJ_DIR):
        if f.endswith(".o") and f not in keep:
            os.remove(os.path.join(OBJ_DIR, f))
    stamp = os.path.join(OBJ_DIR, "link.stamp")
    link_tag = _digest(objs) if all(os.path.exists(o) for o in objs) else ""
    old = open(stamp).read() if os.path.exists(stamp) else ""
    if force or jobs or not os.path.exists(OUT) or old != link_tag:
        _run([NVCC, *ARCH, "-shared", "-cudart", "shared", "-o", OUT, *objs,
              "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-Wno-deprecated-gpu-targets"])
        with open(stamp, "w") as f:
            f.write(link_tag)
    return OUT


def build_host_emu(force: bool = False) -> str:
    """CPU emulation harness used by the non-GPU tests (tests/host_emu)."""
    d = os.path.join(ROOT, "tests", "host_emu")
    out = os.path.join(d, "libymt3_emu.so")
    srcs = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))
    tag = _digest(srcs + _headers())
    stamp = os.path.join(OBJ_DIR, "emu.stamp")
    os.makedirs(OBJ_DIR, exist_ok=True)
    old = open(stamp).read() if os.path.exists(stamp) else ""
    if force or not os.path.exists(out) or old != tag:
        _run([NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
              "-I", os.path.join(ROOT, "include"), "-o", out, *srcs])
        with open(stamp, "w") as f:
            f.write(tag)
    return out


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_emu(force="--force" in sys.argv))
