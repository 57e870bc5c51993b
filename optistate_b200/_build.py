"""In-tree build of the native code: liboptistate_kf.so (nvcc, sm_100a, C ABI) and _optistate_torch.so
(g++, PyTorch C++ extension that links against it).  Both land in optistate_b200/_lib/ so that they travel
with the source tree; nothing is installed into site-packages or a JIT cache.

    python -m optistate_b200._build [--force] [-v]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
KF_LIB = os.path.join(LIB_DIR, "liboptistate_kf.so")
EXT_NAME = "_optistate_torch"
EXT_LIB = os.path.join(LIB_DIR, EXT_NAME + ".so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(LIB_DIR, "obj")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return cand


def build_kf_lib(force: bool = False, verbose: bool = False) -> str:
    """Compiles every csrc/*.cu to an object (in parallel: the kernel instantiations are spread over several
    translation units for exactly that reason) and links them into liboptistate_kf.so."""
    from concurrent.futures import ThreadPoolExecutor

    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(ROOT, "include", "optistate_kf.h"))
    units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    if not force and _newer(KF_LIB, headers + [os.path.join(CSRC, u) for u in units]):
        return KF_LIB  # the library alone is enough (the objects do not travel to the GPU box)
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = [os.path.join(OBJ_DIR, u[:-3] + ".o") for u in units]
    todo = [(u, o) for u, o in zip(units, objs) if force or not _newer(o, headers + [os.path.join(CSRC, u)])]

    def compile_one(uo):
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", "-o", uo[1], os.path.join(CSRC, uo[0])]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as pool:
            list(pool.map(compile_one, todo))
    stale = [o for o in os.listdir(OBJ_DIR) if os.path.join(OBJ_DIR, o) not in objs]
    for o in stale:
        os.remove(os.path.join(OBJ_DIR, o))
    if todo or stale or not _newer(KF_LIB, objs):
        subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", KF_LIB, *objs])
    return KF_LIB


def build_torch_ext(force: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension

    src = os.path.join(CSRC, "torch_binding.cpp")
    if not force and _newer(EXT_LIB, [src, KF_LIB, os.path.join(ROOT, "include", "optistate_kf.h")]):
        return EXT_LIB
    try:
        inc = list(cpp_extension.include_paths("cuda"))
    except TypeError:
        inc = list(cpp_extension.include_paths(True))
    inc += [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = [
        "g++", "-O2", "-fPIC", "-shared", "-std=c++17", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
        f"-DTORCH_EXTENSION_NAME={EXT_NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
        *[f"-I{p}" for p in inc], src, "-o", EXT_LIB,
        f"-L{LIB_DIR}", "-loptistate_kf", f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
        "-ltorch_python", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}",
    ]
    subprocess.check_call(cmd)
    return EXT_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_kf_lib(force, verbose), build_torch_ext(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
