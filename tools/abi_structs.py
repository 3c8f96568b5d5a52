"""Parse the plain-C config structs out of include/vmm_b200.h and render the ctypes mirror a maintainer of the reference
would paste.  tests/test_abi.py checks that functional.EMConfig / HeadConfig AND the snippet in
INTEGRATION.md agree field-for-field with the header.

    python tools/abi_structs.py            # prints the snippet
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTYPES = {"int": "C.c_int", "unsigned": "C.c_uint", "float": "C.c_float", "unsigned long long": "C.c_ulonglong",
          "long long": "C.c_longlong"}


def parse_struct(name, header=None):
    """-> [(c type, field name), ...] of `typedef struct name {...} name;` (scalar fields only)."""
    text = open(header or os.path.join(ROOT, "include", "vmm_b200.h")).read()
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        mm = re.match(r"(unsigned long long|long long|unsigned|int|float) (.+)$", decl)
        assert mm, decl
        for n in mm.group(2).split(","):
            fields.append((mm.group(1), n.strip()))
    return fields


def render(cls, cname):
    fields = parse_struct(cname)
    lines, cur = [], "    _fields_ = ["
    for i, (t, n) in enumerate(fields):
        item = '("%s", %s)' % (n, CTYPES[t]) + (", " if i + 1 < len(fields) else "]")
        if len(cur) + len(item) > 118:
            lines.append(cur.rstrip())
            cur = "                "
        cur += item
    lines.append(cur)
    return "class %s(C.Structure):            # %s (include/vmm_b200.h)\n%s" % (cls, cname, "\n".join(lines))


def snippet():
    return render("EMConfig", "vmm_em_config") + "\n\n" + render("HeadConfig", "vmm_head_config")


if __name__ == "__main__":
    print(snippet())
