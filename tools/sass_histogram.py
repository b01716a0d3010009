#!/usr/bin/env python
"""SASS opcode histogram per hot kernel of libtealeaf_b200.so (cuobjdump -sass; no GPU needed).
usage: python tools/sass_histogram.py > profiles/rNN_sass_histograms.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tealeaf.jl_b200", "csrc", "libtealeaf_b200.so")
HOT = ["k_cg_fused_w_ringILi2ELi3ELi2E", "k_cg_fused_w_ringILi0ELi3ELi3E", "k_cg_fused_w_ringILi2ELi4ELi2E", "k_cg_fused_w_ringILi0ELi4ELi2E", "k_cg_fused_w_ringILi1ELi3ELi3E", "k_cg_fused_w_tmaILb1ELi4ELi2E", "k_cg_fused_r9CgBParams",
       "k_cheby_fused_ringILb0ELi3ELi3E", "k_cheby_pair_ringILi4ELi2ELb0E", "k_cheby_pair_ringILi4ELi2ELb1E",
       "k_ppcg_inner_ringILi3ELi3E", "k_ppcg_pair_ringILi4ELi2ELb0E", "k_ppcg_pair_ringILi4ELi2ELb1E", "k_jacobi_fused_ringILi3ELi3E"]
GROUPS = [("FP64 (DADD/DMUL/DFMA)", r"^(DADD|DMUL|DFMA)$"), ("global/shared data (LDG/STG/LDS/STS/LDGSTS/UTMALDG/UBLKCP)", r"^(LDG|STG|LDS|STS|LDGSTS|UTMALDG|UBLKCP|LDGDEPBAR|DEPBAR)$"),
          ("integer / address (IMAD/IADD3/LEA/VIADD/SHF/LOP3/...)", r"^(IMAD|IADD3|IADD|LEA|VIADD|SHF|LOP3|MOV|UMOV|ULEA|UIADD3|UIMAD|R2UR|S2R|S2UR|CS2R|LDC|LDCU|ULDC|SEL|PRMT|IABS|I2F|F2I|I2FP|UISETP|USEL|ULOP3|USHF)$"),
          ("predicates / control (ISETP/PLOP3/FSEL/BRA/BSSY/BSYNC/...)", r"^(ISETP|PLOP3|FSEL|FSETP|DSETP|BRA|BSSY|BSYNC|EXIT|WARPSYNC|ENDCOLLECTIVE|CALL|RET|NOP|BREAK|YIELD|VOTE|VOTEU|ELECT|P2R|R2P)$"),
          ("shuffles (SHFL)", r"^SHFL$"), ("sync / fences / atomics (BAR/MEMBAR/SYNCS/ATOM/RED/...)", r"^(BAR|MEMBAR|SYNCS|ATOM|ATOMG|ATOMS|RED|ERRBAR|CGAERRBAR|CCTL|FENCE|ACQBULK|UCGABAR_ARV|UCGABAR_WAIT)$")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)[1:]
    print("# SASS opcode histograms of the hot kernels (`cuobjdump -sass`, sm_100a, whole kernel incl. prologue and tail)\n")
    print("Static instruction counts of the compiled kernels (not executed counts): they prove which data-movement instructions the")
    print("kernels use (LDGSTS = cp.async, UTMALDG = TMA tensor copy) and show the instruction mix the ncu issue-slot numbers come from.\n")
    for key in HOT:
        for f in funcs:
            name = f.split("\n", 1)[0].strip()
            if key not in name:
                continue
            ops = collections.Counter()
            for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", f):
                ops[m.group(1)] += 1
            total = sum(ops.values())
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
            print(f"## `{demangled}`  ({total} instructions)\n")
            print("| group | count | share |\n|---|---|---|")
            rest = dict(ops)
            for label, pat in GROUPS:
                n = sum(v for k, v in ops.items() if re.match(pat, k))
                for k in list(rest):
                    if re.match(pat, k):
                        rest.pop(k)
                print(f"| {label} | {n} | {100 * n / total:.0f}% |")
            print(f"| other | {sum(rest.values())} | {100 * sum(rest.values()) / total:.0f}% |\n")
            print("top opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)) + "\n")
            break


if __name__ == "__main__":
    main()
