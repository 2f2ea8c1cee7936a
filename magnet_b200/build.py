"""Build libmagnet_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m magnet_b200.build [--force]

The .so lands in magnet_b200/lib/ (git-ignored, travels to the GPU box with the snapshot).
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# developer variants (e.g. MGB_NVCC_EXTRA=-DMGB_TIMELINE MGB_VARIANT=tl) build into their own directory
_VARIANT = os.environ.get("MGB_VARIANT", "")
LIB_DIR = os.path.join(_HERE, "lib" + ("_" + _VARIANT if _VARIANT else ""))
LIB_PATH = os.path.join(LIB_DIR, "libmagnet_b200.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("MGB_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(_HERE), "include", "magnet_b200.h")]
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.sha256")
    digest = _digest(srcs + headers)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    hdr_digest = _digest(headers)
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        ostamp = obj + ".sha256"
        d = _digest([src]) + hdr_digest
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read().strip() == d:
            continue
        log = open(obj + ".log", "w")
        procs.append((src, obj, ostamp, d, log,
                      subprocess.Popen([nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj], stdout=log, stderr=subprocess.STDOUT)))
    failed = []
    for src, obj, ostamp, d, log, p in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            failed.append((src, open(obj + ".log").read()))
        else:
            open(ostamp, "w").write(d)
            if verbose:
                print(open(obj + ".log").read())
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"--- {s}\n{l}" for s, l in failed))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    open(stamp, "w").write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
