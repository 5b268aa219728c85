"""Builds muax_b200/libmzsearch.so in-tree with nvcc for sm_100a (no torch involved: the library is plain
CUDA runtime + the C ABI of include/mzsearch.h).  One object per translation unit, compiled in parallel and
cached under csrc/_obj/ (stale objects are detected through the header mtimes)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.environ.get("MZ_LIB_OUT") or os.path.join(PKG, "libmzsearch.so")
OBJ_DIR = os.path.join(HERE, "_obj")
INC = [os.path.join(ROOT, "include", f) for f in ("mz_math.h", "mzsearch.h")]


def _h(*names):
    return [os.path.join(HERE, f) for f in names] + INC


# translation unit -> the headers it includes (a TU is rebuilt when it or one of these is newer than its object)
UNITS = {
    "mzsearch.cu": _h("mz_device.cuh", "mz_warp.cuh", "mz_treewarp.cuh", "mz_resident.cuh", "mz_recurrent_tc.cuh"),
    "mz_warp.cu": _h("mz_device.cuh", "mz_warp.cuh"),
    "mz_treewarp.cu": _h("mz_device.cuh", "mz_records.cuh", "mz_resident.cuh", "mz_treewarp.cuh"),
    "mz_resident.cu": _h("mz_device.cuh", "mz_records.cuh", "mz_resident.cuh"),
    "mz_recurrent_tc.cu": _h("mz_device.cuh", "mz_recurrent_tc.cuh"),
}
HEADERS = sorted({p for deps in UNITS.values() for p in deps})
SOURCES = [os.path.join(HERE, u) for u in UNITS]
DEPS = SOURCES + HEADERS
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    return os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(p) for p in DEPS)


def _defines():
    return ["-D" + d for d in os.environ.get("MZ_NVCC_DEFINES", "").split()]


def _compile(unit, force, verbose):
    src = os.path.join(HERE, unit)
    tag = "".join(sorted(os.environ.get("MZ_NVCC_DEFINES", "").split()))
    obj = os.path.join(OBJ_DIR, unit.replace(".cu", (("." + tag) if tag else "") + ".o"))
    deps = [src] + UNITS[unit]
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(p) for p in deps):
        return obj, ""
    cmd = [nvcc_path()] + NVCC_FLAGS + _defines() + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed compiling {unit}")
    return obj, res.stdout + res.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    # several processes (torchrun ranks, xdist workers) may get here at once: one builds, the others wait for the
    # lock and then find the library up to date; the link goes to a temporary name and is renamed into place
    import fcntl
    with open(os.path.join(OBJ_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and up_to_date():
            return OUT
        with ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
            results = list(pool.map(lambda u: _compile(u, force, verbose), UNITS))
        objs = [o for o, _ in results]
        tmp = f"{OUT}.{os.getpid()}.tmp"
        res = subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "--cudart=static",
                              "-Xcompiler", "-fPIC", "-o", tmp] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking libmzsearch.so")
        os.replace(tmp, OUT)
    if verbose:
        print("".join(log for _, log in results))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
