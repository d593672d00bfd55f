"""Build the C-ABI shared library (CUDA kernels + cuFFT glue) in-tree with nvcc.

    python -m ocelot_b200.build [--force] [--verbose]

Produces ``ocelot_b200/libocelot_sc.so`` for sm_100a.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libocelot_sc.so")
SOURCES = ["sc_kernels.cu", "sc_fft.cu", "sc_beam.cu", "sc_lsc.cu", "sc_abi.cu"]
HEADERS = ["sc_device.cuh", "sc_kernels.h", "sc_special.h", os.path.join("..", "..", "include", "ocelot_sc.h")]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; set NVCC or install the CUDA toolkit")
    return cand


def have_nvcc() -> bool:
    try:
        nvcc_path()
        return True
    except RuntimeError:
        return False


def cuda_lib_dir(nvcc: str) -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "lib64")


def command(verbose: bool = False) -> list[str]:
    nvcc = nvcc_path()
    cmd = [nvcc, "-O3", "-std=c++17",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-lineinfo",
           # IEEE division / square root everywhere; FMA contraction is allowed except where the
           # source uses explicit __dmul_rn/__dadd_rn (Green's function, see sc_kernels.cu)
           "-prec-div=true", "-prec-sqrt=true",
           "-Xcompiler", "-fPIC", "-shared",
           "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lcufft", "-Xlinker", "-rpath=" + cuda_lib_dir(nvcc)]
    return cmd


STAMP = LIB + ".srchash"


def source_hash() -> str:
    """Content hash of everything the library is built from (sources, headers, this recipe)."""
    import hashlib
    h = hashlib.sha256()
    for d in [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]:
        with open(d, "rb") as f:
            h.update(os.path.basename(d).encode() + b"\0" + f.read())
    return h.hexdigest()


def up_to_date() -> bool:
    """True when the library was built from the sources as they are now.  Decided by content (a stamp file written
    next to the library), not by modification times: a snapshot copied to another machine keeps the contents, not
    necessarily the mtimes.  Without a stamp, fall back to comparing mtimes."""
    if not os.path.exists(LIB):
        return False
    if os.path.exists(STAMP):
        with open(STAMP) as f:
            return f.read().strip() == source_hash()
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = command(verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
