"""C++ host side: the header-only facade (include/gem_b200/elevation_map.hpp) and the
source-level shim re-exporting the reference's nine entry points (compat/gpu_process_shim.cpp).
CPU: both compile and link against libgem_b200.so (the shim against the stand-in Eigen header,
the real Eigen is not in this image).  GPU: the program runs and both paths agree."""
import os
import subprocess

import pytest

from gem_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cxx", "facade_smoke")


def compile_program():
    lib = build.build()
    srcs = [os.path.join(ROOT, "tests", "cxx", "facade_smoke.cpp"), os.path.join(ROOT, "compat", "gpu_process_shim.cpp")]
    deps = srcs + [lib, os.path.join(ROOT, "include", "gem_b200.h"), os.path.join(ROOT, "include", "gem_b200", "elevation_map.hpp")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return EXE
    cmd = ["g++", "-O2", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "mini_eigen"),
           "-o", EXE] + srcs + ["-L", os.path.dirname(lib), "-lgem_b200", "-Wl,-rpath," + os.path.dirname(lib)]
    subprocess.run(cmd, check=True)
    return EXE


def test_facade_and_shim_compile_and_link():
    exe = compile_program()
    assert os.path.exists(exe)
    out = subprocess.run(["nm", "-C", "--defined-only", exe], capture_output=True, text=True).stdout
    for sym in ("Init_GPU_elevationmap(int, float, float, float)", "Raytracing(int)", "Map_closeloop(float*, float, int, float)",
                "Mapvar_update(int, float)", "Map_optmove(float*, float, float, int, float*)"):
        assert sym in out, sym


@pytest.mark.gpu
def test_facade_and_shim_run_and_agree():
    exe = compile_program()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches=0" in r.stdout
