"""Build the C-ABI CUDA library in-tree (nvcc, sm_100a only).

    python -c "import __graft_entry__ as g; g.build()"      # or: python <pkg>/_build.py [--force] [-v]

Every csrc/*.cu is compiled to an object of its own (in parallel, cached by a digest of the source, the headers and the
flags) and the objects are linked into <pkg>/libupd_b200.so.  The .so is git-ignored (history stays source-only) but is
NOT gpurun-ignored: it travels to the GPU box with the snapshot, which has no reason to compile anything.

    python <pkg>/_build.py --variant NAME -DUPD_X=1 ...      # experimental build -> <pkg>/build/NAME/libupd_b200.so
                                                              # (load it with UPD_LIB_PATH=...; profiles/tools use this)
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_NAME = "libupd_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
BUILD_DIR = os.path.join(PKG_DIR, "build")
SOURCES = ["api.cu", "sampler_simt.cu", "sampler_tc.cu", "sampler_ws.cu", "mpv_reduce.cu", "sigma_est.cu",
           "infill_steps.cu", "stg_steps.cu", "stg_tcn_mma.cu", "fx_fused.cu", "gemm3.cu", "fx_attention.cu", "dts_attention.cu", "dts_attention_tc.cu", "dts_norm.cu"]
HEADERS = ["upd_common.cuh", "sampler_params.cuh", "sampler_math.cuh", "sampler_epi.cuh", "tc_helpers.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # fusion only where the source says fmaf(): keeps the reference's fp32 rounding
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def _header_digest(extra):
    h = hashlib.sha256()
    for name in HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    with open(os.path.join(INCLUDE, "upd_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h


def source_digest(extra=()):
    h = _header_digest(extra)
    for name in SOURCES:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile_one(name, obj_dir, extra, force, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(obj_dir, name[:-3] + ".o")
    h = _header_digest(extra)
    with open(src, "rb") as f:
        h.update(f.read())
    digest, stamp = h.hexdigest(), obj + ".digest"
    if not force and os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + ["-I", INCLUDE, "-c", "-o", obj, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed on {}:\n{}{}".format(name, proc.stdout, proc.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return obj, proc.stderr


def build_library(force=False, verbose=False, variant=None, extra=()):
    """Compile csrc/*.cu into <pkg>/libupd_b200.so (or build/<variant>/libupd_b200.so) unless up to date."""
    obj_dir = os.path.join(BUILD_DIR, variant or "release")
    lib_path = os.path.join(obj_dir, LIB_NAME) if variant else LIB_PATH
    os.makedirs(obj_dir, exist_ok=True)
    stamp = lib_path + ".digest"
    digest = source_digest(extra)
    if not force and os.path.exists(lib_path) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return lib_path
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, obj_dir, extra, force, verbose), SOURCES))
    if verbose:
        for (obj, log), name in zip(results, SOURCES):
            print("==== " + name)
            print(log)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_path] + [o for o, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return lib_path


SELFTEST_LIB = os.path.join(PKG_DIR, "libupd_selftest.so")


def build_selftest(force=False):
    """tests-only library with the tcgen05 descriptor known-answer kernel (csrc/selftest_umma.cu); not part of the
    product ABI and never loaded by the package."""
    src = os.path.join(CSRC, "selftest_umma.cu")
    h = _header_digest(())
    with open(src, "rb") as f:
        h.update(f.read())
    digest, stamp = h.hexdigest(), SELFTEST_LIB + ".digest"
    if not force and os.path.exists(SELFTEST_LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return SELFTEST_LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-I", INCLUDE, "-o", SELFTEST_LIB, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed on selftest_umma.cu:\n" + proc.stdout + proc.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return SELFTEST_LIB


if __name__ == "__main__":
    args = sys.argv[1:]
    variant = None
    if "--variant" in args:
        i = args.index("--variant")
        variant = args[i + 1]
        del args[i:i + 2]
    print(build_library(force="--force" in args, verbose="-v" in args, variant=variant,
                        extra=[a for a in args if a.startswith("-D")]))
    if variant is None:
        print(build_selftest(force="--force" in args))
