"""Builds kmerpapa_b200/libkpapa.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_PKG, "csrc")
SOURCES = ["kp_api.cu", "kp_plan.cpp"]
HEADERS = ["kp_kernels.cuh", "kp_fiber.cuh", "kp_math.cuh", "kp_tables.h", "kp_plan.h", "kp_log_data.h", "kp_shard_owners.h"]
OUT = os.path.join(_PKG, "libkpapa.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # no implicit FMA contraction: the score arithmetic must round like the CPU reference
    "-Xcompiler", "-fPIC",
    "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(_SRC, f) for f in SOURCES + HEADERS] + [os.path.join(_PKG, "..", "include", "kmerpapa_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    if not force and not is_stale():
        return OUT
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    env = dict(os.environ)
    # the image's CC/CXX point at a repackaged gcc; let nvcc use the system host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, cwd=_SRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return OUT


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
