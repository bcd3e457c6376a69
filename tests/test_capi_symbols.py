"""CPU: the C-ABI library loads and exports every entry point the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "jess_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from pyjess_b200 import _capi
    assert sorted(_capi.EXPORTS) == declared_functions()


def test_library_exports_every_symbol():
    from pyjess_b200 import _build, _capi
    _build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name
    lib.jb_abi_version.restype = ctypes.c_int
    assert lib.jb_abi_version() == 3


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the engine must fail loudly, never compute on the CPU."""
    from pyjess_b200 import _capi
    lib = _capi.load_library()
    if lib.jb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.EngineError) as err:
        _capi.Context(0)
    assert "no CPU fallback" in str(err.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyjess_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "libjess_ref" not in text, f
