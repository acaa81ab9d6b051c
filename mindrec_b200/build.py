"""Build libmindrec_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python -m mindrec_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MREC_BUILD_SUFFIX / MREC_BUILD_DEFINES: an experimental variant next to the product library, e.g.
#   MREC_BUILD_SUFFIX=_t16 MREC_BUILD_DEFINES=-DMREC_SEG_TILE=16 python -m mindrec_b200.build
# (run it with MREC_LIB_PATH=mindrec_b200/libmindrec_b200_t16.so); without them this is the product build.
_SUFFIX = os.environ.get("MREC_BUILD_SUFFIX", "")
OBJ = os.path.join(HERE, "build" + _SUFFIX)
LIB = os.path.join(HERE, "libmindrec_b200%s.so" % _SUFFIX)

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("MREC_BUILD_DEFINES", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for p in [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha1"
    dig = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile_one, srcs))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    if verbose and log:
        print(log)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
