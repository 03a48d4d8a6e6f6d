"""Builds libxenomapper_b200.so (sm_100a only) and the synthetic-data helper in-tree.

    python -m xenomapper_b200.build
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libxenomapper_b200.so")
SOURCES = ["xm_kernels.cu", "xm_api.cu"]
HEADERS = ["xm_common.h", "xm_parse.h", "xm_tile.h", "xm_walk.h", "xm_launch.h", "xm_stream.h", "xm_bam.h", "xm_scan2.cuh", "xm_emit.cuh", "xm_shard.h", "xm_nccl.h", "xm_headers.h", "xm_bgzf.h", "xm_inflate.h", "xm_bamchain.h", "xm_deflate.h", "xm_fmtg.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-lz", "-ldl"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "xenomapper_b200.h")]
    if force or _stale(LIB, deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout)
        if r.returncode:
            raise RuntimeError("nvcc failed building libxenomapper_b200.so")
        with open(os.path.join(HERE, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    # test twin whose QNAME hash depends on the name's length only (every pair of equally long names collides): the GPU
    # tests load it through XM_LIB_PATH to show that the byte compare behind the hash is what decides (never the default)
    weak = os.path.join(HERE, "libxenomapper_b200_weakhash.so")
    if force or _stale(weak, deps):
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DXM_WEAK_HASH", "-o", weak] + [os.path.join(CSRC, f) for f in SOURCES]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout)
            raise RuntimeError("nvcc failed building the weak-hash test twin")
    if os.environ.get("XM_BUILD_PROF"):
        # profiling twin with per-phase cycle counters (never loaded by the package; XM_LIB_PATH selects it in bench.py)
        prof = os.path.join(HERE, "libxenomapper_b200_prof.so")
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + NVCC_FLAGS + ["-DXM_PHASE_TIMING=1", "-o", prof] + [os.path.join(CSRC, f) for f in SOURCES]
        subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
    from . import synth
    synth_so = os.path.join(HERE, "_xm_synth.so")
    if force or _stale(synth_so, [os.path.join(CSRC, "xm_synth.c")]):
        synth.build()
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
