"""Build libwvd.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m video_styler_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The resulting ``video_styler_b200/libwvd.so`` is git-ignored but travels to
the GPU box with the repo snapshot.  cudart is linked statically; the driver API (cuTensorMapEncodeTiled) is
resolved at run time, so the library also loads on a CPU-only box (for the symbol-export tests).
"""
import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libwvd.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "elementwise.cu", "gemm_sm100.cu", "attention_sm100.cu", "attention_pair_sm100.cu", "attention_cg2_sm100.cu", "attention_cg2p_sm100.cu", "t5_attention.cu", "fallthrough_f32.cu"]
HEADERS = [os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "softmax_math.cuh"), os.path.join(CSRC, "host_utils.h"), os.path.join(ROOT, "include", "wvd.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# developer builds, e.g. WVD_NVCC_FLAGS=-DWVD_ATTN_PROF (in-kernel cycle counters of the attention softmax)
NVCC_FLAGS += os.environ.get("WVD_NVCC_FLAGS", "").split()


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def source_hash() -> str:
    """sha256 over every file the library is compiled from (csrc/* and include/wvd.h), in sorted order: baked into
    the library (wvd_build_info) so that a prebuilt libwvd.so can be checked against the sources it travels with."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    for f in files + [os.path.join(ROOT, "include", "wvd.h")]:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = nvcc_path()
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    all_src = [os.path.join(CSRC, x) for x in SOURCES]
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        deps = [s] + HEADERS + [os.path.abspath(__file__)]
        extra = []
        if src == "api.cu":          # carries the source hash of the whole library: rebuilt whenever any source changes
            deps += all_src
            extra = [f'-DWVD_SOURCE_HASH="{source_hash()}"']
        if force or _stale(o, deps):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for src, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
