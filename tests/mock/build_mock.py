"""Builds tests/mock/_build/libmock_host.so = host_mirror.cpp (the PRODUCT's host mirror, unchanged) + mock_it.c (a CPU
stand-in for the device library, answered by the oracle).  Test infrastructure: lets the CPU suite run the mirror's
file-to-file calls.  Returns a ctypes.CDLL with the lowdin_host_* argtypes of openlowdin_b200.capi applied."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libmock_host.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(ROOT, "openlowdin_b200", "csrc", "host_mirror.cpp"), os.path.join(HERE, "mock_it.c"),
            os.path.join(ROOT, "oracle", "it_oracle.c"), os.path.join(ROOT, "oracle", "blas_shim.c")]
    if os.path.exists(OUT) and all(os.path.getmtime(s) < os.path.getmtime(OUT) for s in srcs):
        return OUT
    objs = []
    for s in srcs:
        o = os.path.join(os.path.dirname(OUT), os.path.basename(s) + ".o")
        cc = ["/usr/bin/g++", "-std=c++17"] if s.endswith(".cpp") else ["/usr/bin/gcc", "-std=c11", "-fopenmp"]
        subprocess.run(cc + ["-O1", "-fPIC", "-c", s, "-o", o], check=True)
        objs.append(o)
    subprocess.run(["/usr/bin/g++", "-shared", "-o", OUT] + objs + ["-lgomp", "-lm", "-ldl"], check=True)
    return OUT


def load():
    L = C.CDLL(build())
    return L


ERI_OUT = os.path.join(HERE, "_build", "libmock_eri.so")


def build_eri():
    """The product's ERI evaluator (it_eri.cuh, __host__ __device__) compiled for the host by nvcc: CPU check of the device algorithm."""
    os.makedirs(os.path.dirname(ERI_OUT), exist_ok=True)
    src = os.path.join(HERE, "eri_host.cu")
    deps = [src, os.path.join(ROOT, "openlowdin_b200", "csrc", "it_eri.cuh"), os.path.join(ROOT, "openlowdin_b200", "csrc", "it_kernels.cuh")]
    if os.path.exists(ERI_OUT) and all(os.path.getmtime(s) < os.path.getmtime(ERI_OUT) for s in deps):
        return ERI_OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-fopenmp",
                    "-ccbin", "/usr/bin/g++", "-o", ERI_OUT, src, "-lgomp"], check=True)
    return ERI_OUT
