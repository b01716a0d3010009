"""julia/TeaLeafB200.jl cannot be executed here (no Julia toolchain), so it is checked statically:
every `ccall` must name a symbol the header declares, with the argument list (count and C type
class) of the matching prototype in include/tealeaf_b200.h and of the ctypes table the executed
tests use; the SolveInfo struct must mirror tl_solve_info field by field."""
import ctypes as C
import os
import re

from conftest import ROOT

JL = open(os.path.join(ROOT, "julia", "TeaLeafB200.jl"), encoding="utf-8").read()
HDR = open(os.path.join(ROOT, "include", "tealeaf_b200.h")).read()


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def jl_class(t):
    t = t.strip()
    if t in ("Cint",):
        return "int"
    if t == "Cuint":
        return "uint"
    if t == "Clong":
        return "long"
    if t == "Clonglong":
        return "longlong"
    if t == "Cdouble":
        return "double"
    if t in ("Cvoid", "Nothing"):
        return "void"
    if t == "Cstring" or t.startswith("Ptr{") or t.startswith("Ref{"):
        return "ptr"
    raise AssertionError(f"unmapped Julia C type {t!r}")


def c_class(t):
    t = re.sub(r"\bconst\b", "", t).strip()
    if "*" in t:
        return "ptr"
    base = t.split()[:-1] if re.search(r"\w+\s+\w+$", t) else t.split()
    base = " ".join(base) or t
    return {"int": "int", "unsigned": "uint", "long": "long", "long long": "longlong", "double": "double",
            "void": "void"}[base]


def lp64(cls):
    """ctypes aliases c_longlong to c_long on LP64: compare those two as one class"""
    return "i64" if cls in ("long", "longlong") else cls


def ctypes_class(t):
    if t is None:
        return "void"
    if t in (C.c_int,):
        return "int"
    if t is C.c_uint:
        return "uint"
    if t is C.c_long:
        return "long"
    if t is C.c_longlong:
        return "longlong"
    if t is C.c_double:
        return "double"
    return "ptr"


def header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|const char \*)\s*(tl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argl = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
        protos[name] = ("ptr" if "*" in ret else c_class(ret + " x"), [c_class(a) for a in argl])
    return protos


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\),\s*(\w+(?:\{[^()]*?\})?),\s*\((.*?)\),", JL, flags=re.S):
        name, ret, args = m.group(1), m.group(2), m.group(3)
        calls.append((name, jl_class(ret), [jl_class(a) for a in split_top(args)]))
    return calls


def test_every_ccall_matches_the_header_and_the_ctypes_table():
    from tealeaf_jl_b200 import lib
    protos = header_prototypes()
    assert sorted(protos) == sorted(lib.SIGNATURES)
    calls = julia_ccalls()
    assert len(calls) >= 20
    for name, ret, args in calls:
        assert name in protos, f"{name} is not declared in include/tealeaf_b200.h"
        hret, hargs = protos[name]
        assert (ret, args) == (hret, hargs), (name, (ret, args), (hret, hargs))
        cret, cargs = lib.SIGNATURES[name]
        assert (lp64(ret), [lp64(a) for a in args]) == (lp64(ctypes_class(cret)), [lp64(ctypes_class(a)) for a in cargs]), name


def test_julia_binds_every_entry_point_a_host_needs():
    bound = {c[0] for c in julia_ccalls()}
    needed = {"tl_create", "tl_destroy", "tl_last_error", "tl_set_field", "tl_get_field", "tl_copy_field", "tl_halo_update",
              "tl_cg_init", "tl_cg_calc_w", "tl_cg_calc_ur", "tl_cg_calc_p", "tl_copy_u", "tl_calc_residual", "tl_finalise",
              "tl_solve_finished", "tl_field_summary", "tl_cg_solve", "tl_cheby_solve", "tl_ppcg_solve", "tl_jacobi_solve",
              "tl_paint_states"}
    assert needed <= bound, needed - bound


def test_solveinfo_struct_mirrors_tl_solve_info():
    from tealeaf_jl_b200 import lib
    m = re.search(r"struct SolveInfo\n(.*?)\nend", JL, flags=re.S)
    fields = [f.strip() for f in re.split(r"[;\n]", m.group(1)) if f.strip()]
    jl = [(f.split("::")[0], jl_class(f.split("::")[1])) for f in fields]
    hm = re.search(r"typedef struct tl_solve_info \{(.*?)\} tl_solve_info;", HDR, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", hm.group(1), flags=re.S)
    hdr = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        ctype, names = decl.rsplit(" ", 1)[0], decl
        toks = decl.replace(",", " ").split()
        tname = "long long" if toks[:2] == ["long", "long"] else toks[0]
        for n in toks[2 if tname == "long long" else 1:]:
            hdr.append((n, c_class(tname + " x")))
    assert jl == hdr, (jl, hdr)
    assert [(n, lp64(c)) for n, c in jl] == [(n, lp64(ctypes_class(t))) for n, t in lib.SolveInfo._fields_]


def _strip_julia(text):
    """drop docstrings / strings / comments (identifiers inside them are not code)"""
    text = re.sub(r'"""(.|\n)*?"""', '""', text)
    text = re.sub(r'"(\\.|[^"\\\n])*"', '""', text)
    return "\n".join(line.split("#", 1)[0] for line in text.splitlines())


def test_no_unexported_reference_name_is_used_unqualified():
    """`using TeaLeaf` only brings the reference's EXPORTED names into scope (src/chunk.jl:4-7, src/settings.jl:4-8,
    src/TeaLeaf.jl:2); anything else the binding takes from the reference must be written `TeaLeaf.name` (or imported
    explicitly), otherwise the first call throws UndefVarError -- which no test here could notice, Julia being absent.
    The name lists come from the reference tree (tests/golden/make_reference_names.py)."""
    import json
    names = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_names.json"), encoding="utf-8"))
    in_scope = set(names["using_TeaLeaf"])
    code = _strip_julia(JL)
    in_scope |= {n.strip() for m in re.finditer(r"using\s+TeaLeaf(?:\.\w+)*\s*:\s*(.*)", code) for n in m.group(1).split(",")}
    in_scope |= {n.strip() for m in re.finditer(r"using\s+\.\.TeaLeafB200\s*:\s*(.*)", code) for n in m.group(1).split(",")}
    # names the binding defines itself (functions, structs, modules, consts) -- unqualified definitions only
    local = set(re.findall(r"^\s*function\s+([A-Za-z_][\w!]*)\s*\(", code, flags=re.M))
    local |= set(re.findall(r"^\s*(?:mutable\s+)?struct\s+([A-Za-z_]\w*)", code, flags=re.M))
    local |= set(re.findall(r"^\s*module\s+([A-Za-z_]\w*)", code, flags=re.M))
    local |= set(re.findall(r"^\s*const\s+([A-Za-z_]\w*)", code, flags=re.M))
    local |= set(re.findall(r"^\s*([A-Za-z_][\w!]*)\([^=\n]*\)\s*=(?!=)", code, flags=re.M))
    offenders = []
    for m in re.finditer(r"(?<![\w.:!])([A-Za-z_][\w!]*)", code):
        name = m.group(1)
        if name in names["defined"] and name not in in_scope and name not in local:
            line = code[:m.start()].count("\n") + 1
            offenders.append((name, line))
    assert not offenders, f"non-exported reference names used unqualified (write TeaLeaf.<name>): {offenders}"
    # and the qualified uses must name things the reference really defines
    for m in re.finditer(r"TeaLeaf\.(?:Kernels\.|CG\.|Cheby\.|PPCG\.|Jacobi\.)?([A-Za-z_][\w!]*)", code):
        if m.group(1) not in ("Kernels", "TeaLeafB200"):
            assert m.group(1) in names["defined"], f"TeaLeaf.{m.group(1)} does not exist in the reference"


def test_reference_names_fixture_is_current():
    """When the reference tree is present (this container), the committed fixture must match it."""
    import json
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/src"):
        import pytest
        pytest.skip("reference tree not present (GPU box)")
    path = os.path.join(ROOT, "tests", "golden", "reference_names.json")
    before = json.load(open(path, encoding="utf-8"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_reference_names.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert json.load(open(path, encoding="utf-8")) == before
