"""In-tree build of the CUDA library (sm_100a) — `python -m sdfibm_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsdfibm_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "sdfibm_cuda.cu"), os.path.join(HERE, "csrc", "mesh_host.cpp")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "device_math.cuh"), os.path.join(HERE, "csrc", "interact_kernels.cuh"), os.path.join(ROOT, "include", "sdfibm_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",               # strict predicates: no FMA contraction
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-shared",
    "-ldl",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """`defines` / `out`: experiment builds (kernel variants selected by -D macros into another file; SDFIBM_B200_LIB loads it)."""
    if not force and out == LIB and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


MESH_LIB = os.path.join(HERE, "libsdfibm_mesh.so")


def build_mesh(force: bool = False) -> str:
    """The Foam-free mesh helpers alone (g++, no CUDA): what sdfibm_b200.mesh loads, so that building a mesh never maps the
    CUDA library (the CPU reference arm of bench.py times the reference's code with only this helper loaded)."""
    src = os.path.join(HERE, "csrc", "mesh_host.cpp")
    hdr = os.path.join(ROOT, "include", "sdfibm_b200.h")
    if force or not os.path.exists(MESH_LIB) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(MESH_LIB):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DSDFIBM_MESH_STANDALONE", "-o", MESH_LIB, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("mesh helper build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return MESH_LIB


HOST_LIB = os.path.join(HERE, "libsdfibm_host.so")
HOST_DIR = os.path.join(HERE, "host")
HOST_SOURCES = [os.path.join(HOST_DIR, "solidcloud.cpp"), os.path.join(HOST_DIR, "capi_host.cpp")]
RUNNER = os.path.join(HERE, "sdfibm_b200_run")
VOF_RUNNER = os.path.join(HERE, "sdfibm_b200_vof")


def _host_deps():
    deps = [os.path.join(ROOT, "include", "sdfibm_b200_host.h"), os.path.join(ROOT, "include", "sdfibm_b200.h"),
            os.path.join(HERE, "csrc", "device_math.cuh"), LIB]
    for d, _, files in os.walk(HOST_DIR):
        deps += [os.path.join(d, f) for f in files if f.endswith((".h", ".cpp", ".H"))]
    return deps


def build_host(force: bool = False) -> str:
    """Host façade (C++17, g++): libsdfibm_host.so + the stand-alone runner, linked against the CUDA library."""
    outs = [HOST_LIB, RUNNER, VOF_RUNNER]
    if not force and all(os.path.exists(o) for o in outs):
        t = min(os.path.getmtime(o) for o in outs)
        if all(os.path.getmtime(d) <= t for d in _host_deps()):
            return HOST_LIB
    common = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-Wall"]
    link = ["-L" + HERE, "-lsdfibm_b200", "-Wl,-rpath,$ORIGIN"]
    for cmd in (common + ["-shared", "-o", HOST_LIB] + HOST_SOURCES + link,
                common + ["-o", RUNNER, os.path.join(HOST_DIR, "standalone_main.cpp")] + HOST_SOURCES + link,
                common + ["-o", VOF_RUNNER, os.path.join(HOST_DIR, "tool_vof", "vof_main.cpp"), os.path.join(HOST_DIR, "solidcloud.cpp")] + link):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return HOST_LIB


def build_oracle(force: bool = False) -> str:
    d = os.path.join(ROOT, "oracle")
    lib = os.path.join(d, "liboracle.so")
    src = os.path.join(d, "oracle.cpp")
    hdr = os.path.join(ROOT, "include", "sdfibm_b200.h")
    if force or not os.path.exists(lib) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(lib):
        r = subprocess.run(["make", "-C", d, "-B", "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return lib


def build_reference_oracle(force: bool = False, reference: str = "/root/reference"):
    """oracle/_ref/libsdfibm_ref.so: the reference's own hot-path translation units, compiled unmodified from where they lie
    (only where the reference tree exists — this container, not the GPU box).  Test infrastructure; see oracle/Makefile."""
    if not os.path.isdir(os.path.join(reference, "src")):
        return None
    d = os.path.join(ROOT, "oracle")
    lib = os.path.join(d, "_ref", "libsdfibm_ref.so")
    deps = [os.path.join(d, "refshim", f) for f in ("ref_bridge.cpp", "foam_shim.h")] + [os.path.join(HERE, "host", "foamlite.h"), os.path.join(d, "Makefile")]
    if force or not os.path.exists(lib) or max(os.path.getmtime(x) for x in deps) > os.path.getmtime(lib):
        r = subprocess.run(["make", "-C", d, "-B", "ref", f"REFERENCE={reference}"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference oracle build failed:\n" + r.stdout + r.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_mesh(force="--force" in sys.argv))
    print(build_host(force="--force" in sys.argv))
    print(build_oracle(force="--force" in sys.argv))
    print(build_reference_oracle(force="--force" in sys.argv))
