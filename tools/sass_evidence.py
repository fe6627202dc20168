"""SASS evidence of the Blackwell-native data path: per kernel, the count of bulk-async-copy (TMA unit) and mbarrier instructions
next to the classic ones.  `python tools/sass_evidence.py > profiles/r0N_sass_tma.txt` (needs cuobjdump, no GPU)."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "harp_b200", "libharpb200.so")
MNEMONICS = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "ATOMS", "REDUX", "RED.E", "ATOMG", "FFMA", "DFMA", "HMMA", "UTC")

txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
tot = collections.Counter()
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter()
    for m in MNEMONICS:
        n = len(re.findall(r"\b" + re.escape(m), f))
        if n:
            c[m] = n
            tot[m] += n
    rows.append((name, c))
demangle = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("cuobjdump -sass harp_b200/libharpb200.so  (sm_100a)")
print("library totals:", dict(tot))
print("\nkernels with bulk async copies (cp.async.bulk -> UBLKCP) and mbarrier traffic (SYNCS):")
for (name, c), d in zip(rows, demangle):
    if c.get("UBLKCP") or c.get("SYNCS"):
        print(f"  {d[:150]}\n      {dict(c)}")
print("\nall kernels:")
for (name, c), d in zip(rows, demangle):
    print(f"  {d[:120]:120s} {dict(c)}")
m = re.search(r"UBLKCP[^\n]*", txt)
if m:
    i = txt.index(m.group(0))
    print("\nexcerpt around the first bulk copy:\n" + "\n".join(l[:110] for l in txt[max(0, i - 900):i + 400].splitlines()))
