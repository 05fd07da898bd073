"""Static SASS mnemonic counts per kernel of the built library (cuobjdump -sass), written to profiles/r02_sass_summary.txt.
Runs on the CPU box: it is the committed evidence that the hot kernels are tcgen05 / TMA / TMEM code and which packed forms they use."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "world-in-world_b200", "libwiw_b200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r02_sass_summary.txt")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
COLS = [("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
        ("UTMASTG", r"\bUTMASTG"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("USETMAXREG", r"\bUSETMAXREG"), ("FFMA2/FMUL2/FADD2", r"\bF(FMA|MUL|ADD)2\b"),
        ("FMNMX3", r"\bFMNMX3"), ("HMMA", r"\bHMMA"), ("STL/LDL", r"\b(STL|LDL)")]
kern, rows = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("wiw::", "")
        kern = name
        rows[kern] = {"mufu": collections.Counter(), **{c: 0 for c, _ in COLS}}
        continue
    if kern is None or "/*" not in line:
        continue
    ins = line.split("*/", 1)[-1]
    for c, pat in COLS:
        if re.search(pat, ins):
            rows[kern][c] += 1
    mm = re.search(r"\bMUFU\.(\w+)", ins)
    if mm:
        rows[kern]["mufu"][mm.group(1)] += 1
with open(OUT, "w") as f:
    f.write("SASS mnemonics per kernel of world-in-world_b200/libwiw_b200.so (cuobjdump -sass, sm_100a; static instruction counts; scripts/sass_summary.py)\n")
    f.write("tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2) | tcgen05.ld/st -> LDTM/STTM | TMA -> UTMALDG/UTMASTG | tcgen05.commit -> UTCBAR | setmaxnreg -> USETMAXREG | "
            "f32x2 -> FFMA2/FMUL2/FADD2 | 3-input max -> FMNMX3 | mma.sync -> HMMA | STL/LDL = register spills\n\n")
    f.write(f"{'kernel':<44}" + "".join(f"{c:>19}" for c, _ in COLS) + "   MUFU ops\n")
    for k, r in rows.items():
        f.write(f"{k:<44}" + "".join(f"{r[c]:>19}" for c, _ in COLS) + "   " + ", ".join(f"MUFU.{a}:{b}" for a, b in sorted(r["mufu"].items())) + "\n")
print(open(OUT).read()[:200])
