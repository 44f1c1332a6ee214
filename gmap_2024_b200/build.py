"""Builds the in-tree native libraries (sm_100a only; nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libgmapdp_b200.so")
# test build: the genome bridge's per-lane tie lists hold ONE entry, so that their overflow path (the fills
# repeated with every tie resolved on the spot) runs on ordinary boxes; loaded only by the GPU parity tests
LIB_TIECAP1 = os.path.join(CSRC, "libgmapdp_b200_tiecap1.so")
SOURCES = ["gmapdp_kernels.cu", "gmapdp_shim.cpp", "gmapdp_stream.cpp", "gmapchain_kernels.cu", "gmapchain_shim.cpp"]
HEADERS = ["gmapdp_layout.h", "gmapdp_tables.h", "gmapdp_internal.h", "gmapdp_genome.h", "../../include/gmapdp_b200.h", "../../include/gmapdp_shim.h",
           "../../include/gmapchain_b200.h", "../../include/gmapdp_stream.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_native(force=False, verbose=False):
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    nvcc = os.environ.get("NVCC", "nvcc")
    if force or _stale(LIB, deps):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        subprocess.check_call(cmd, cwd=CSRC)
    if force or _stale(LIB_TIECAP1, deps):
        subprocess.check_call([nvcc] + NVCC_FLAGS + ["-DGEN_TIECAP=1", "-o", LIB_TIECAP1] + SOURCES, cwd=CSRC)
    return LIB
