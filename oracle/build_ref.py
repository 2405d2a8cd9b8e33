"""TEST INFRASTRUCTURE — builds the reference's own CPU path into oracle/_ref/.

Compiles, with /usr/bin/g++ directly (never the reference's setup.py/Makefile), the
"cpu" source set the reference lists in setup.py:224-240 plus pybind/minkowski.cpp,
from the sources where they lie under /root/reference, against the torch headers in
this image. Outputs go ONLY to oracle/_ref/ (git-ignored, shipped to the GPU box):

    oracle/_ref/MinkowskiEngineBackend/_C.so      the compiled reference CPU backend
    oracle/_ref/MinkowskiEngine/                  the reference's Python package, as
                                                  `pip install --target` would place it
    oracle/_ref/obj/*.o                           object files (incremental rebuilds)

Toolchain trap: the image exports CXX=/opt/gcc/bin/g++, which
links libstdc++ statically and makes every C++ exception/iostream inside an extension
crash. We therefore call /usr/bin/g++ explicitly and verify `ldd` afterwards.

Usage:  python oracle/build_ref.py [--jobs 8] [--force]
No-op (returns False) if /root/reference is absent (the GPU box): the prebuilt files
under oracle/_ref/ are used as they are.
"""
import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MEB200_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CXX = "/usr/bin/g++"

CPU_SOURCES = [  # reference setup.py:224-240 ("cpu" set) + pybind/minkowski.cpp
    "src/math_functions_cpu.cpp",
    "src/coordinate_map_manager.cpp",
    "src/convolution_cpu.cpp",
    "src/convolution_transpose_cpu.cpp",
    "src/local_pooling_cpu.cpp",
    "src/local_pooling_transpose_cpu.cpp",
    "src/global_pooling_cpu.cpp",
    "src/broadcast_cpu.cpp",
    "src/pruning_cpu.cpp",
    "src/interpolation_cpu.cpp",
    "src/quantization.cpp",
    "src/direct_max_pool.cpp",
    "pybind/minkowski.cpp",
]


def so_path():
    return os.path.join(OUT, "MinkowskiEngineBackend", "_C.so")


def available():
    return os.path.isfile(so_path()) and os.path.isdir(os.path.join(OUT, "MinkowskiEngine"))


def _flags():
    import pybind11
    import torch
    from torch.utils.cpp_extension import include_paths

    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = [os.path.join(HERE, "ref_shim"), os.path.join(REF, "src"),
           os.path.join(REF, "src", "3rdparty")]
    inc += include_paths() + [pybind11.get_include(), sysconfig.get_paths()["include"]]
    cflags = ["-O3", "-fopenmp", "-fPIC", "-std=c++17", "-DCPU_ONLY", "-w",
              "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cflags += ["-I" + p for p in inc]
    ldflags = ["-shared", "-fopenmp", "-L" + torch_lib, "-Wl,-rpath," + torch_lib,
               "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    return cflags, ldflags


def _compile(src, obj, cflags):
    if os.path.isfile(obj) and os.path.getmtime(obj) > os.path.getmtime(src):
        return obj
    subprocess.run([CXX, *cflags, "-c", src, "-o", obj], check=True)
    return obj


def build(jobs=8, force=False, verbose=True):
    if not os.path.isdir(REF):
        return False
    if available() and not force:
        return True
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(os.path.dirname(so_path()), exist_ok=True)
    cflags, ldflags = _flags()
    todo = [(os.path.join(REF, s), os.path.join(objdir, os.path.basename(s) + ".o"))
            for s in CPU_SOURCES]
    if verbose:
        print(f"[oracle/_ref] compiling {len(todo)} reference sources with {CXX} (-j{jobs})",
              flush=True)
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = list(ex.map(lambda so: _compile(so[0], so[1], cflags), todo))
    subprocess.run([CXX, *objs, *ldflags, "-o", so_path()], check=True)
    open(os.path.join(OUT, "MinkowskiEngineBackend", "__init__.py"), "w").close()
    ldd = subprocess.run(["ldd", so_path()], capture_output=True, text=True).stdout
    assert "libstdc++.so.6" in ldd and "libgomp.so.1" in ldd, \
        "oracle/_ref/_C.so must link libstdc++/libgomp dynamically:\n" + ldd
    # the reference's Python package, placed as an install would place it (not tracked)
    dst = os.path.join(OUT, "MinkowskiEngine")
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(os.path.join(REF, "MinkowskiEngine"), dst,
                    ignore=shutil.ignore_patterns("__pycache__"))
    if verbose:
        print("[oracle/_ref] built", so_path(), flush=True)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.jobs, a.force)
    print("reference present:", ok, "| oracle/_ref available:", available())
    sys.exit(0)
