"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vmm_b200.h")).read()
    return sorted(set(re.findall(r"VMM_API\s+[\w\s\*]+?\b(vmm_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from multi_view_sort_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 8
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/vmm_b200.h but not exported: {missing}"
    lib.vmm_version.restype = ctypes.c_int
    assert lib.vmm_version() >= 1            # host-only call, no GPU needed


def test_no_oracle_import_in_product():
    """The product package must never reach into oracle/ (the checker is not the product)."""
    pkg = os.path.join(ROOT, "multi_view_sort_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
