"""The C-ABI shared library loads on a box without a GPU and exports every symbol include/*.h declares.
No compute call is made here (no GPU); device-less calls must fail loudly, never fall back to a CPU path."""
import ctypes
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(lowdin_(?:it|host)_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("lowdin_it_create", "lowdin_it_set_species", "lowdin_it_ao_push_stacks", "lowdin_it_transform",
                 "lowdin_it_download_pairs", "lowdin_it_download_quads", "lowdin_it_transform_all",
                 "lowdin_it_transform_inter_all", "lowdin_it_comm_init"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from openlowdin_b200 import capi
    L = capi.load()
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/ but not exported by liblowdin_itgpu.so"
    assert sorted(capi.ABI_SYMBOLS + capi.HOST_SYMBOLS) == declared_symbols()


def test_library_is_sm100a_cuda_code():
    from openlowdin_b200 import capi
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", capi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_device_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from openlowdin_b200 import capi
    L = capi.load()
    h = ctypes.c_void_p()
    assert L.lowdin_it_create(0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in L.lowdin_it_last_error(None) or b"CPU" in L.lowdin_it_last_error(None)
    with pytest.raises(capi.LowdinITError):
        capi.Transformer(0)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under openlowdin_b200/ or include/ may reference it."""
    for path in glob.glob(os.path.join(ROOT, "openlowdin_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
            src = open(path).read()
            assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, path


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: both headers compile as C99 with a C compiler (no C++ types, no torch), and a C program
    that references every declared function links against the library."""
    syms = declared_symbols()
    src = tmp_path / "abi_check.c"
    body = "\n".join(f"  p[{i}] = (fn)&{s};" for i, s in enumerate(syms))
    src.write_text('#include "lowdin_it.h"\n#include "lowdin_it_host.h"\n#include <stdio.h>\n'
                   f"typedef void (*fn)(void);\nint main(void) {{\n  fn p[{len(syms)}];\n{body}\n  printf(\"%d\\n\", (int)(sizeof p / sizeof p[0]));\n  return p[0] == 0;\n}}\n")
    from openlowdin_b200 import capi
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(capi.lib_path())
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", libdir, "-llowdin_itgpu", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and int(r.stdout) == len(syms)
