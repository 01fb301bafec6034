"""
Build the CUDA library in-tree:  thetis_b200/libthetis_b200.so  (sm_100a only).

    python -m thetis_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libthetis_b200.so")
SOURCES = ["tb_kernels.cu", "tb_tracer.cu", "tb_api.cu"]
HEADERS = [os.path.join(CSRC, "tb_internal.h"), os.path.join(CSRC, "tb_device.cuh"), os.path.join(CSRC, "tb_wd_mass.cuh"),
           os.path.join(HERE, "..", "include", "thetis_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas=-v",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu of the package for sm_100a into one shared library."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    # developer A/B builds: TB_EXTRA_NVCC="-DTB_T_MINB=5" THETIS_B200_LIB=/path/variant.so python -m thetis_b200.build --force
    extra = os.environ.get("TB_EXTRA_NVCC", "").split()
    out = os.environ.get("THETIS_B200_LIB") or LIB
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libthetis_b200.so")
    return out


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
    print(LIB)
