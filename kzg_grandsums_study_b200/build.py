"""In-tree build of libkzgb200.so (hand-written sm_100a CUDA + the C ABI of include/kzgb200.h).

`python -m kzg_grandsums_study_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a
GPU; the .so lands next to this file so that it travels to the GPU box with the tree.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libkzgb200.so")
SOURCES = ["core.cu", "host.cu", "frops.cu", "ntt.cu", "argument.cu", "msm.cu", "srs.cu", "prover.cu", "mgpu.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
if os.environ.get("KZGB200_DEBUG_BOUNDS") == "1":
    # debug build: every hand-computed index of the MSM kernels is checked (msm.cu KZG_IDX_OK); slower, lanes serialised
    NVCC_FLAGS.append("-DKZG_BOUNDS_CHECK")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps_digest(extra):
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(extra)
    return h.hexdigest()


def _compile_one(src):
    os.makedirs(OBJ_DIR, exist_ok=True)
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    stamp = obj + ".sha"
    with open(path, "rb") as f:
        digest = _deps_digest(f.read())
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return src, 0, "up to date"
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", path, "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(obj + ".log", "w") as f:
        f.write(p.stdout)
    if p.returncode == 0:
        with open(stamp, "w") as f:
            f.write(digest)
    return src, p.returncode, p.stdout


def _tree_digest():
    h = hashlib.sha256()
    for src in SOURCES:
        with open(os.path.join(CSRC, src), "rb") as f:
            h.update(f.read())
    h.update(_deps_digest(b"").encode())
    return h.hexdigest()


def source_digest(names):
    """sha256 over the named csrc files: the key under which tools/summarize_profiles.py files ncu-counted figures
    (profiles/traffic.json) and bench.py looks them up -- a kernel edit silently invalidates nothing"""
    h = hashlib.sha256()
    for name in names:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


MSM_SOURCES = ["msm.cu", "field.cuh", "ec.cuh", "frops.cu"]
NTT_SOURCES = ["ntt.cu", "field.cuh"]


def build(verbose=False):
    """Compile every translation unit (in parallel) and link libkzgb200.so.  Returns the library path.
    A stamp next to the library (it travels to the GPU box with it) makes an unchanged tree a no-op."""
    stamp = LIB_PATH + ".stamp"
    digest = _tree_digest()
    if os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        if verbose:
            print("[build] libkzgb200.so is up to date")
        return LIB_PATH
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(_compile_one, SOURCES))
    relink = not os.path.exists(LIB_PATH)
    for src, rc, out in results:
        if rc != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if out != "up to date":
            relink = True
        if verbose:
            print("[build] %s: %s" % (src, "ok" if out != "up to date" else out))
    if relink:
        objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose="-q" not in sys.argv))
