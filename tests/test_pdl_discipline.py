"""CPU: every kernel that is launched with programmatic stream serialization (launch_pdl,
csrc/common.cuh) must contain `griddepcontrol.wait` (SASS: ACQBULK).  All hot kernels call
`griddepcontrol.launch_dependents` at entry, so a dependent kernel WITHOUT the wait could start
reading its inputs before the kernel that produces them has finished (the
halo conv kernel had the trigger but not the wait).  Checked on the compiled objects."""

import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "smallpebble_b200", "csrc")
OBJ = os.path.join(ROOT, "smallpebble_b200", "_lib", "obj")


def _kernels_launched_with_pdl(source: str):
    names = set()
    for m in re.finditer(r"launch_pdl\(\s*([A-Za-z_][A-Za-z_0-9]*)", source):
        name = m.group(1)
        if name == "kern":
            # `auto kern = some_kernel<...>;` a few lines above the launch
            before = source[: m.start()]
            defs = re.findall(r"\bkern\s*=\s*([A-Za-z_][A-Za-z_0-9]*)", before)
            assert defs, "launch_pdl(kern, ...) without a visible definition of kern"
            name = defs[-1]
        names.add(name)
    return names


def _sass_functions(obj_path: str):
    """{mangled function name: set of opcode mnemonics we care about}"""
    out = subprocess.run(["cuobjdump", "-sass", obj_path], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), set())
            continue
        if cur is not None:
            for op in ("ACQBULK", "PREEXIT"):
                if re.search(rf"\b{op}\b", line):
                    cur.add(op)
    return funcs


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_every_pdl_launched_kernel_waits_for_its_prerequisite_grid():
    from smallpebble_b200 import build

    build.build()
    checked = 0
    for fname in sorted(os.listdir(CSRC)):
        if not fname.endswith(".cu"):
            continue
        with open(os.path.join(CSRC, fname)) as fh:
            launched = _kernels_launched_with_pdl(fh.read())
        if not launched:
            continue
        funcs = _sass_functions(os.path.join(OBJ, fname[:-3] + ".o"))
        for kernel in sorted(launched):
            matches = {f: ops for f, ops in funcs.items() if re.search(rf"\d+{kernel}(I|E|P|v|$)", f)}
            assert matches, f"{fname}: no compiled function found for kernel {kernel}"
            for f, ops in matches.items():
                assert "ACQBULK" in ops, (
                    f"{fname}: {kernel} ({f}) is launched with launch_pdl but has no "
                    f"griddepcontrol.wait")
                checked += 1
    assert checked >= 30


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_kernels_that_trigger_dependents_also_wait():
    """The converse discipline: a kernel that releases its dependents early (PREEXIT) and never
    waits would break the transitive stream order the buffer free list relies on."""
    from smallpebble_b200 import build

    build.build()
    for fname in sorted(os.listdir(OBJ)):
        if not fname.endswith(".o"):
            continue
        for f, ops in _sass_functions(os.path.join(OBJ, fname)).items():
            if "PREEXIT" in ops:
                assert "ACQBULK" in ops, f"{fname}: {f} triggers dependents but never waits"
