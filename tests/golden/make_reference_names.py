#!/usr/bin/env python
"""Generates tests/golden/reference_names.json from the reference tree (run HERE, where /root/reference exists; the
fixture travels): the names `using TeaLeaf` brings into scope (the `export` lines of src/TeaLeaf.jl, src/settings.jl,
src/chunk.jl), the names `using TeaLeaf.Kernels` would add (@exportAll: everything src/kernels.jl defines), and every
name the reference defines at all.  tests/test_julia_binding.py uses it to check that julia/TeaLeafB200.jl never uses
a non-exported reference name unqualified (it cannot be executed here: no Julia toolchain)."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_names.json")

DEF_PATTERNS = [r"^\s*function\s+([A-Za-z_][\w!]*)", r"^\s*([A-Za-z_][\w!]*)\([^=]*\)\s*=(?!=)", r"^\s*const\s+([A-Za-z_]\w*)",
                r"^\s*(?:@with_kw\s+)?(?:mutable\s+)?struct\s+([A-Za-z_]\w*)", r"^\s*module\s+([A-Za-z_]\w*)"]


def defined(path):
    names = set()
    for line in open(path, encoding="utf-8"):
        for pat in DEF_PATTERNS:
            m = re.match(pat, line)
            if m:
                names.add(m.group(1))
        m = re.match(r"^\s*@enum\s+(\w+)\s+(.*)", line)
        if m:
            names.add(m.group(1))
            names.update(m.group(2).split())
    return names


def exported(path):
    names = set()
    for line in open(path, encoding="utf-8"):
        m = re.match(r"^\s*export\s+(.*)", line)
        if m:
            names.update(n.strip() for n in m.group(1).split(",") if n.strip())
    return names


src = os.path.join(REF, "src")
top = ["TeaLeaf.jl", "settings.jl", "chunk.jl"]
out = {
    "generated_by": "tests/golden/make_reference_names.py from /root/reference (Laura7089/TeaLeaf.jl @ e696c54)",
    "using_TeaLeaf": sorted(set().union(*(exported(os.path.join(src, f)) for f in top)) | {"TeaLeaf"}),
    "using_TeaLeaf_Kernels": sorted(defined(os.path.join(src, "kernels.jl")) - {"Kernels"}),
    "defined": sorted(set().union(*(defined(os.path.join(dp, f)) for dp, _, fs in os.walk(src) for f in fs if f.endswith(".jl")))),
}
json.dump(out, open(OUT, "w"), indent=1, ensure_ascii=False)
print(OUT, {k: len(v) for k, v in out.items() if isinstance(v, list)})
