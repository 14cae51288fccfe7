"""CPU: the C-ABI shared library loads and exports every symbol include/*.h declares; the
ctypes binding covers exactly that set; the product package never touches the oracle."""

import ctypes
import os
import re

from smallpebble_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smallpebble_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:int|size_t|const char\*)\s+(sp_\w+)\s*\(", text, flags=re.M)
    assert len(names) > 30
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.sp_abi_version() == 1


def test_binding_matches_header():
    assert sorted(_abi.SIGNATURES) == declared_functions()


def test_conv_geom_struct_layout():
    assert ctypes.sizeof(_abi.ConvGeom) == 13 * 4
    text = open(HEADER).read()
    body = re.search(r"typedef struct sp_conv_geom \{(.*?)\}", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r"int32_t([^;]*);", body) for f in decl.split(",")]
    assert fields == [n for n, _ in _abi.ConvGeom._fields_]


def test_error_reporting_without_device():
    lib = _abi.load()
    rc = lib.sp_gemm(0, 0, -1, 1, 1, None, 1, 0, None, 1, 0, None, 1, 0, 1, 0, None)
    assert rc == -1 and b"negative extent" in lib.sp_last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under smallpebble_b200/ (nor the C sources)
    may import, call, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "smallpebble_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "numpy_path" not in text, f
                assert "/root/reference" not in text, f
