"""Build liblowdin_itgpu.so in-tree with nvcc for sm_100a (B200).  No CPU fallback is built."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblowdin_itgpu.so")
SOURCES = ["it_api.cu", "host_mirror.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lowdin_it.h"),
                                                                  os.path.join(HERE, "..", "include", "lowdin_it_host.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-ccbin", "/usr/bin/g++", "-o", LIB] + srcs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building liblowdin_itgpu.so")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:  # registers / spills / smem per kernel (tracked in git)
        f.write("".join(l for l in r.stderr.splitlines(True) if "Compile time" not in l))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
