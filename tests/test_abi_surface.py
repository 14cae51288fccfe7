"""CPU: the C-ABI shared library loads and exports every symbol include/*.h declares; the
ctypes binding covers exactly that set; the product package never touches the oracle."""

import ctypes
import os
import re

import pytest

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


def test_fastcall_binding_matches_ctypes_without_device():
    """csrc/fastcall.c: the CPython binding calls the same exports as ctypes -- argument
    conversion (ints, None, floats, NumPy integer scalars, ctypes buffers), the status check and
    the error text are exercised on calls that fail or finish before touching a device."""
    import numpy as np

    if os.environ.get("SP_NO_FASTCALL") == "1":
        pytest.skip("the fast binding is switched off by SP_NO_FASTCALL=1")
    _abi.load()
    fast = _abi._load_fast()
    assert fast is not None, "smallpebble_b200/_lib/_sp_fastcall.so was not built"
    fn = _abi.f.sp_gemm
    assert type(fn).__name__ == "FastFn" and fn.__name__ == "sp_gemm"
    with __import__("pytest").raises(_abi.SmallPebbleCudaError, match="negative extent"):
        fn(0, 0, -1, 1, 1, None, 1, 0, None, 1, 0, None, 1, 0, 1, 0, None)
    # a NumPy integer scalar passes its VALUE (not the address of its buffer)
    with __import__("pytest").raises(_abi.SmallPebbleCudaError, match="negative extent"):
        fn(0, 0, np.int64(-1), 1, 1, None, 1, 0, None, 1, 0, None, 1, 0, 1, 0, None)
    # wrong arity is a TypeError, not a crash
    with __import__("pytest").raises(TypeError):
        fn(0, 0)
    # M == 0: returns before any launch; counts as a launching call of the binding
    before = _abi.counter.launches
    fn(0, 0, 0, 4, 4, None, 4, 0, None, 4, 0, None, 4, 0, 1, 0, None)
    assert _abi.counter.launches == before + 1
    # a ctypes object is passed by address: the flag is written through it
    flag = ctypes.c_int(7)
    _abi.f.sp_host_is_pinned(np.zeros(4).ctypes.data, flag)
    assert flag.value == 0


def test_ctypes_fallback_binding(monkeypatch):
    """SP_NO_FASTCALL=1 (or a missing _sp_fastcall.so) leaves the ctypes binding in charge."""
    import subprocess
    import sys

    code = ("import os, sys; sys.path.insert(0, %r); os.environ['SP_NO_FASTCALL'] = '1';"
            "from smallpebble_b200 import _abi; _abi.load();"
            "fn = _abi.f.sp_gemm; assert type(fn).__name__ == 'function', type(fn);\n"
            "try:\n"
            "    fn(0, 0, -1, 1, 1, None, 1, 0, None, 1, 0, None, 1, 0, 1, 0, None)\n"
            "except _abi.SmallPebbleCudaError as e:\n"
            "    assert 'negative extent' in str(e); print('ok')\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-500:]
