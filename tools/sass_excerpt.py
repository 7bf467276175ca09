#!/usr/bin/env python
"""SASS evidence for profiles/: instruction mix of the two K3+K4 kernels in kmerpapa_b200/libkpapa.so (sm_100a cubin) and
excerpts of their hot loops.  python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt   (needs cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "kmerpapa_b200", "libkpapa.so")
KERNELS = {"rows": "_Z17kp_dp_rows_kernelILi15ELb0ELi232ELi0EEv10KpDpParams", "fiber": "_Z18kp_dp_fiber_kernelILb0EEv13KpFiberParams"}


def sass(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, SO], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;?\s*/\*", line)
        if m:
            ins.append((m.group(1), m.group(2).rstrip(" ;")))
    return ins


def mix(ins):
    c = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        c[".".join(op.split(".")[:3]) if op.startswith(("LDG", "STG", "LDS", "STS", "MUFU", "UBLKCP", "UBLKPF", "UTMA", "LDGSTS")) else op.split(".")[0]] += 1
    return c


def excerpt(ins, first, n):
    i0 = next(i for i, (_, t) in enumerate(ins) if first(t))
    return "\n".join(f"    /*{a}*/  {t}" for a, t in ins[max(0, i0 - 4): i0 + n])


print("# SASS of the K3+K4 kernels of kmerpapa_b200/libkpapa.so (cuobjdump -sass, sm_100a), made by tools/sass_excerpt.py")
for name, fun in KERNELS.items():
    ins = sass(fun)
    c = mix(ins)
    print(f"\n## kp_dp_{name}_kernel: {len(ins)} instructions (with its out-of-line device functions)")
    keys = ["LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "FADD2", "FFMA2", "FMUL2", "FMNMX", "FMNMX3", "FADD", "DFMA", "DADD", "DMUL",
            "DSETP", "MUFU.RCP64H", "MUFU.LG2", "MUFU.RCP", "I2F", "F2F", "BAR", "CALL", "UBLKPF", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "STL", "LDL"]
    print("   ", ", ".join(f"{k} {sum(v for kk, v in c.items() if kk == k or kk.startswith(k + '.'))}" for k in keys))
    if name == "rows":
        print("\n### child-tile stream (phase D): eight LDG.128 per step with immediate group offsets, packed FADD2, FMNMX")
        print(excerpt(ins, lambda t: "LDG.E.128" in t, 64))
        print("\n### fast self-score (kp_self_score_fast): reciprocal seed + two Newton steps, table log, Horner log(1-p), error bound")
        print(excerpt(ins, lambda t: "MUFU.RCP64H" in t, 96))
    else:
        print("\n### fiber stream: batches of 2 x 4 x 2 predicated LDG.128, packed FADD2, FMNMX")
        print(excerpt(ins, lambda t: "LDG.E.128.CONSTANT" in t, 72))
