#!/usr/bin/env python3
"""SASS evidence per kernel of libcgx_b200.so: counts of the mnemonics that prove the Blackwell-native
mechanisms (TMA bulk-tensor copies, cp.async.bulk copies, mbarrier transactions) next to the fp64 arithmetic.
    python tools/sass_summary.py > profiles/sass_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "new_cg_variants_b200", "_obj")
PAT = {"UTMALDG": r"UTMALDG", "UBLKCP": r"UBLKCP", "SYNCS (mbarrier)": r"SYNCS\.", "BAR.SYNC": r"BAR\.SYNC", "LDG.E.128": r"LDG\.E\.128", "STG.E.128": r"STG\.E\.128",
       "LDS.128": r"LDS\.128", "DFMA": r"DFMA", "DMUL": r"DMUL", "DADD": r"DADD", "LDL/STL (spill)": r"\b(LDL|STL)"}
print("# cuobjdump -sass of the objects linked into libcgx_b200.so (sm_100a): static instruction counts per kernel")
print("# kernel | " + " | ".join(PAT))
for obj in sorted(os.listdir(OBJ)):
    if not obj.endswith(".o") or "pers" in obj:
        continue
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[cur] = collections.Counter()
            continue
        if cur:
            for k, p in PAT.items():
                if re.search(p, line):
                    counts[cur][k] += 1
    for k, c in counts.items():
        if any(s in k for s in ("stencil_tma", "pr_fused", "csr_stream", "csr_bulk", "ew_kernel<4", "ew_kernel<3", "ew_kernel<5")) and "<" in k:
            print(f"{obj}: {k[:90]} | " + " | ".join(str(c[p]) for p in PAT))
