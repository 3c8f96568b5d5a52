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


def _tool():
    import importlib.util
    spec = importlib.util.spec_from_file_location("abi_structs", os.path.join(ROOT, "tools", "abi_structs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_config_structs_match_header_field_for_field():
    """vmm_em_config / vmm_head_config parsed out of include/vmm_b200.h == functional.EMConfig / HeadConfig (names, order,
    ctypes types): a host mirror that is a field short hands the library a short struct."""
    import ctypes as C
    from multi_view_sort_b200 import functional as F
    tool = _tool()
    ctype = {"int": C.c_int, "unsigned": C.c_uint, "float": C.c_float, "unsigned long long": C.c_ulonglong,
             "long long": C.c_longlong}
    for cname, cls in (("vmm_em_config", F.EMConfig), ("vmm_head_config", F.HeadConfig)):
        want = [(n, ctype[t]) for t, n in tool.parse_struct(cname)]
        got = [(n, t) for n, t in cls._fields_]
        assert got == want, (cname, got, want)


def test_pointer_structs_match_header():
    """vmm_em_params / vmm_head_params: the pointer tables of the Python mirror list the header's fields in order."""
    from multi_view_sort_b200 import functional as F
    text = open(os.path.join(ROOT, "include", "vmm_b200.h")).read()
    for cname, fields in (("vmm_em_params", F.EM_PARAM_FIELDS), ("vmm_head_params", F.HEAD_PARAM_FIELDS)):
        m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, re.S)
        assert m, cname
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        names = re.findall(r"\*\s*(\w+)\s*[;,]", body)
        assert names == [f for f, _ in fields], (cname, names)


def test_integration_snippet_matches_header():
    """The ctypes stub printed in INTEGRATION.md section 3 is the generated one (tools/abi_structs.py)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"# BEGIN abi_structs[^\n]*\n(.*?)\n# END abi_structs", doc, re.S)
    assert m, "INTEGRATION.md lost its generated struct block"
    assert m.group(1).strip() == _tool().snippet().strip()
