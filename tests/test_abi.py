"""CPU-side checks of the drop-in boundary: the native library loads, exports every symbol the headers
declare, and its structs agree with the reference's GPU build (-DGPU_OPEN) field by field."""
import ctypes as C
import json
import os
import re
import subprocess

import pytest

from helpers import GOLDEN, ROOT
from pangulu_b200 import _lib

VTYPES = ["r64", "r32", "cr64", "cr32"]


@pytest.mark.parametrize("vtype", VTYPES)
def test_library_exports_every_declared_symbol(vtype):
    L = _lib.load(vtype)
    for fn in _lib.PLATFORM_FUNCTIONS:
        assert getattr(L, f"pangulu_platform_{_lib.PLATFORM_ID}_{fn}")
    for fn in _lib.EXTENSION_FUNCTIONS + _lib.DRIVER_FUNCTIONS:
        assert getattr(L, fn)
    assert L.pangulu_sm100_version().decode() == f"pangulu-sm100 {vtype} id={_lib.PLATFORM_ID} sm_100a"


def test_headers_and_loader_agree():
    """Every function declared in include/*.h is exported, and nothing is exported that the headers hide."""
    text = open(os.path.join(ROOT, "include", "pangulu_platform_0201001.h")).read()
    psm = set(re.findall(r"PSM\((\w+)\)\(", text)) - {"fn"}
    assert psm == set(_lib.PLATFORM_FUNCTIONS)
    ext = set(re.findall(r"\b(pangulu_sm100_\w+)\(", text))
    assert ext == set(_lib.EXTENSION_FUNCTIONS)
    drv = open(os.path.join(ROOT, "include", "pangulu_sm100_driver.h")).read()
    names = set(re.findall(r"\b(pangulu_(?:sm100_driver_\w+|sm100_reference_\w+|init|gstrf|gstrs|gssv|finalize))\(", drv))
    assert names == set(_lib.DRIVER_FUNCTIONS)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path("r64")], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for fn in _lib.PLATFORM_FUNCTIONS:
        assert f"pangulu_platform_{_lib.PLATFORM_ID}_{fn}" in exported


def test_the_21_functions_match_the_reference_list():
    """The platform function list is the one the reference's dispatch generator uses (build_helper.py:8-32)."""
    ref = "/root/reference/build_helper.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this box")
    names = re.findall(r'^\s*\["(\w+)",', open(ref).read(), flags=re.M)
    assert names == _lib.PLATFORM_FUNCTIONS


def test_struct_layout_matches_reference_gpu_build():
    with open(os.path.join(GOLDEN, "abi_offsets.json")) as f:
        golden = json.load(f)
    mine = {f"slot.{n}": getattr(_lib.Slot, n).offset for n, _ in _lib.Slot._fields_}
    mine.update({f"task.{n}": getattr(_lib.Task, n).offset for n, _ in _lib.Task._fields_})
    mine["sizeof(slot)"], mine["sizeof(task)"] = C.sizeof(_lib.Slot), C.sizeof(_lib.Task)
    assert mine == golden
    probe = os.path.join(ROOT, "oracle", "_ref", "libpangulu_ref_abi_gpu.so")
    if os.path.exists(probe):  # live check against the reference headers where the probe was built
        lib = C.CDLL(probe)
        off = (C.c_longlong * 30)()
        lib.refabi_offsets(off)
        assert list(off) == list(golden.values())
        assert [lib.refabi_task_ids(i) for i in range(4)] == [_lib.TASK_GETRF, _lib.TASK_TSTRF, _lib.TASK_GESSM, _lib.TASK_SSSSM]


def test_header_static_asserts_compile_as_c_and_cxx(tmp_path):
    """include/*.h are plain C headers (no torch, no C++ types in signatures) and carry the layout asserts."""
    src = tmp_path / "t.c"
    src.write_text('#include "pangulu_sm100_driver.h"\nint main(void){return sizeof(pangulu_storage_slot_t)==144?0:1;}\n')
    inc = os.path.join(ROOT, "include")
    for cc, flags in (("gcc", ["-std=c11"]), ("g++", ["-std=c++17", "-x", "c++"])):
        for vt in ("R64", "CR32"):
            exe = tmp_path / f"t_{cc}_{vt}"
            r = subprocess.run([f"/usr/bin/{cc}", *flags, f"-DCALCULATE_TYPE_{vt}", "-I", inc, str(src), "-o", str(exe)], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            assert subprocess.run([str(exe)]).returncode == 0


def test_no_cpu_fallback_without_device():
    """Without a GPU the numeric path must fail loudly instead of falling back to anything on the CPU."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    code = ("import numpy as np, sys; sys.path.insert(0, %r);"
            "from pangulu_b200 import PanguLU, matrices; PanguLU(matrices.laplace2d(6), nb=8)") % ROOT
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "PanguLU ERROR" in (r.stderr + r.stdout)
    code2 = ("import numpy as np, sys; sys.path.insert(0, %r);"
             "from pangulu_b200 import PanguLU, matrices; s = PanguLU(matrices.laplace2d(6), nb=8, host_only=True); s.gstrf()") % ROOT
    r = subprocess.run([os.sys.executable, "-c", code2], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
