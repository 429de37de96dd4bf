"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from convdr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b2f.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2f_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_symbol_of_the_header():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b2f.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "convdr_b200/_lib.py binds a different set than the header declares"


def test_exported_symbols_are_plain_c_linkage():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in header_symbols():
        assert name in exported


def test_library_is_self_contained_no_torch_no_libcuda_link():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libcuda.so" not in out and "libcudart" not in out


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.b2f_version().decode().startswith("b2f ")
    assert "sm_100a" in lib.b2f_version().decode()


def test_argument_validation_without_device():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.b2f_create(128, None, 0, C.byref(h)) != 0
    assert "768" in _lib.last_error()
    assert lib.b2f_ntotal(None) == -1
    assert lib.b2f_num_shards(None) == 0
    assert lib.b2f_reset(None) != 0


@pytest.mark.skipif(_lib.load().b2f_device_count() > 0, reason="this check is for the GPU-less container")
def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly, not compute on the CPU."""
    import convdr_b200.faiss_compat as faiss
    assert faiss.get_num_gpus() == 0
    index = faiss.IndexFlatIP(768)          # lazy: construction is allowed (the driver builds it first)
    x = np.zeros((4, 768), dtype=np.float32)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        index.add(x)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        index.search(x, 2)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "convdr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
    code = "import sys; import convdr_b200, convdr_b200.faiss_compat, convdr_b200.driver, convdr_b200.dist; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
