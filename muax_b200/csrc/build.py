"""Builds muax_b200/libmzsearch.so in-tree with nvcc for sm_100a (no torch involved: the library is plain
CUDA runtime + the C ABI of include/mzsearch.h)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libmzsearch.so")
SOURCES = [os.path.join(HERE, "mzsearch.cu")]
DEPS = SOURCES + [os.path.join(HERE, f) for f in ("mz_device.cuh", "mz_fused.cuh", "mz_group.cuh", "mz_lane.cuh", "mz_lane2.cuh", "mz_resident.cuh")] + [
    os.path.join(ROOT, "include", f) for f in ("mz_math.h", "mzsearch.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    return os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(p) for p in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
           "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "--cudart=static",
           "-o", OUT] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    for d in os.environ.get("MZ_NVCC_DEFINES", "").split():
        cmd.insert(1, "-D" + d)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmzsearch.so")
    if verbose:
        print(res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
