#!/usr/bin/env python
"""Compile the REFERENCE's own gpu_process.cu (from /root/reference, unmodified, never copied)
against the stand-in Eigen header of oracle/mini_eigen into oracle/_ref/ (git-ignored, travels
to the GPU box).  Two builds: the reference's flags (-O3, default FMA contraction) and a
-fmad=false twin whose arithmetic matches the oracle's definition bit for bit.
TEST INFRASTRUCTURE: only tests/test_reference_pin.py and bench.py's extra reference-kernel
timing load these libraries."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CU = "/root/reference/elevation_mapping/elevation_mapping/cuda/gpu_process.cu"
OUT = os.path.join(HERE, "_ref")


def build():
    if not os.path.exists(REF_CU):
        print("reference source not present; keeping prebuilt oracle/_ref if any")
        return 0
    os.makedirs(OUT, exist_ok=True)
    base = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++14", "-rdc=true", "-shared",
            "-Xcompiler", "-fPIC", "-w", "-I", os.path.join(HERE, "mini_eigen")]
    for name, extra in (("libgpu_ref.so", []), ("libgpu_ref_nofma.so", ["-fmad=false"])):
        out = os.path.join(OUT, name)
        srcs = [REF_CU, os.path.join(HERE, "ref_harness.cu")]
        if os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in srcs):
            continue
        subprocess.run(base + extra + ["-o", out] + srcs, check=True)
        print("built", out)
    return 0


if __name__ == "__main__":
    sys.exit(build())
