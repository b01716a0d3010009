"""Every evidence file the documents cite under profiles/ exists in the tree (names with {a,b} alternatives and * expand)."""
import glob
import os
import re

from conftest import ROOT

DOCS = ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md"),
        os.path.join("profiles", "r02m_ncu_summary.md"), os.path.join("profiles", "r02q_ncu_summary.md"))


def expand(s):
    m = re.search(r"\{([^{}]*)\}", s)
    if not m:
        return [s]
    return [x for alt in m.group(1).split(",") for x in expand(s[:m.start()] + alt + s[m.end():])]


def test_cited_profile_files_exist():
    missing, cited = [], 0
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc), encoding="utf-8").read()
        for m in re.finditer(r"`((?:profiles/)?r0[12][a-z]_[A-Za-z0-9_{},.*\[\]-]+\.(?:log|json|jsonl|csv|md|txt))`", text):
            name = m.group(1) if m.group(1).startswith("profiles/") else "profiles/" + m.group(1)
            for n in expand(name):
                cited += 1
                if not glob.glob(os.path.join(ROOT, n)):
                    missing.append((doc, n))
    assert cited > 100 and not missing, missing
