"""In-tree build of libcrfr.so: every .cu under csrc/ compiled for sm_100a with nvcc, linked into one C-ABI library.

The library sits next to this file so that it travels with the repo snapshot to the GPU box (it is git-ignored, not
gpurun-ignored).  No torch headers are involved: the ABI is plain C (include/crfr.h).
"""
import concurrent.futures as cf
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libcrfr.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libcrfr.so")


def _digest(paths):
    """Content hash of the sources keyed by their path RELATIVE to the repo, so that a snapshot of the repo at another
    absolute location (the GPU box) recognises the library that travelled with it as up to date."""
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.relpath(p, ROOT).encode() + b"\0" + f.read())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, verbose=True):
    """Compiles csrc/*.cu (in parallel) and links libcrfr.so; a no-op when the sources are unchanged."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
        [os.path.join(ROOT, "include", "crfr.h")]
    stamp = LIB + ".stamp"            # next to the library: travels with it (the object directory may not)
    dig = _digest(deps)

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig

    if not force and fresh():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    # one builder at a time (torchrun starts one process per GPU, all of which call build()); the others wait on the
    # lock and then find a fresh library.  The link goes to a temporary name and is renamed into place atomically, so
    # a concurrent loader can never map a half-written file.
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return LIB
        return _build_locked(srcs, stamp, dig, verbose)


def _build_locked(srcs, stamp, dig, verbose):
    nvcc = _nvcc()

    headers = sorted([os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] +
                     [os.path.join(ROOT, "include", "crfr.h")])
    hdig = _digest(headers)

    def one(src):
        # incremental: an object is reused when its own source and every header are unchanged
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        ostamp, odig = obj + ".stamp", _digest([src]) + hdig + " ".join(FLAGS[:4])
        if os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == odig:
            return obj
        cmd = [nvcc] + ARCH + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(ostamp, "w") as f:
            f.write(odig)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, srcs))
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + ARCH + ["-shared", "-o", tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    with open(stamp + ".tmp", "w") as f:
        f.write(dig)
    os.replace(stamp + ".tmp", stamp)
    if verbose:
        print("built", LIB, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
