"""Build the C-ABI CUDA library in-tree (nvcc, sm_100a only).

    python -c "import __graft_entry__ as g; g.build()"      # or: python <pkg>/_build.py

The .so is git-ignored (history stays source-only) but is NOT gpurun-ignored: it travels to the
GPU box with the snapshot, which has no reason to compile anything.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_NAME = "libupd_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
SOURCES = ["api.cu", "sampler_simt.cu", "sampler_tc.cu", "sampler_tc3.cu", "sampler_tc3w.cu", "selftest_umma.cu", "mpv_reduce.cu", "sigma_est.cu",
           "infill_steps.cu", "stg_steps.cu", "fx_fused.cu", "fx_attention.cu", "dts_attention.cu", "dts_norm.cu"]
HEADERS = ["upd_common.cuh", "sampler_params.cuh", "tc_helpers.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # fusion only where the source says fmaf(): keeps the reference's fp32 rounding
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def source_digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    with open(os.path.join(INCLUDE, "upd_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into <pkg>/libupd_b200.so unless an up-to-date build exists."""
    stamp = LIB_PATH + ".digest"
    digest = source_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
