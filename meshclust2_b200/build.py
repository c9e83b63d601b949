"""Build recipe for lib/libmeshclust2_b200.so: hand-written CUDA for sm_100a + the C ABI, in-tree.

nvcc cross-compiles without a GPU.  -fmad=false keeps the fp64 epilogue free of fused multiply-adds so it follows
the reference's expression order (the integer kernels are unaffected).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("mc2_api.cu", "kmer_count.cu", "pair_score.cu", "mean_shift.cu", "text_ingest.cu", "host_encode.cpp")]
HDR = [os.path.join(HERE, "csrc", "mc2_internal.cuh"), os.path.join(HERE, "..", "include", "meshclust2_b200.h")]
OUT = os.path.join(HERE, "lib", "libmeshclust2_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
         "-Xcompiler", "-fPIC,-fopenmp,-O2", "-shared", "-lgomp"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in SRC + HDR + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    print("[meshclust2_b200.build]", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
