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

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "4",
]


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
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "optistate_kf.h"))
    if not force and _newer(KF_LIB, srcs):
        return KF_LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", KF_LIB, os.path.join(CSRC, "kf_abi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
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
