"""Build libprism_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

The library has no PyTorch dependency: `extern "C"`, raw pointers, cudaStream_t.  It is
built next to the package (prism_b200/lib/) so that the .so travels with the source tree;
nvcc cross-compiles sm_100a on a machine without a GPU.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libprism_b200.so")
STAMP = os.path.join(LIB_DIR, "libprism_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # bit-exact paths use explicit _rn intrinsics; keep IEEE div/sqrt and never fast-math
    "-prec-div=true", "-prec-sqrt=true",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(ROOT, "include", "prism_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())      # not the absolute path: the tree is copied to other boxes
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return nvcc if os.path.exists(nvcc) else None


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into lib/libprism_b200.so.  Returns the library path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()

    def fresh():
        if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
            with open(STAMP) as fh:
                return fh.read().strip() == digest
        return False

    if fresh():
        return LIB_PATH
    # one builder at a time (torchrun starts several ranks at once); late-comers find a fresh library
    import fcntl
    lock = open(os.path.join(LIB_DIR, ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if fresh():
            return LIB_PATH
        return _build_locked(digest, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(digest, verbose):
    nvcc = find_nvcc()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # prebuilt library shipped with the tree, no toolchain on this box
        raise RuntimeError("nvcc not found and no prebuilt libprism_b200.so present")
    objs = []
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    tmp = LIB_PATH + ".tmp"
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    os.replace(tmp, LIB_PATH)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
