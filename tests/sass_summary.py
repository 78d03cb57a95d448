"""Writes profiles/r02_sass_summary.md: per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics in the built library
(cuobjdump -sass detectorch_b200/csrc/libdetectorch_b200.so) + the library's dynamic dependencies."""
import collections, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "detectorch_b200", "csrc", "libdetectorch_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern = None
C = collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        C[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    c = C[kern]
    c["instr"] += 1
    base = op.split(".")[0]
    if base == "UTCHMMA":
        c["UTCHMMA"] += 1
        if ".2CTA" in op:
            c["2CTA"] += 1
    elif base.startswith("UTC") and base.endswith("MMA"):
        c["UTCother"] += 1
    elif base in ("UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "F2FP", "SYNCS", "BAR"):
        c[base] += 1
lines = ["# SASS evidence (cuobjdump -sass detectorch_b200/csrc/libdetectorch_b200.so, final round-2 build; regenerate with `python tests/sass_summary.py`)", "",
         "Per-kernel counts of the tcgen05 / TMEM / TMA mnemonics (B200_PROFILING.md: UTCHMMA = tcgen05.mma (the kind::f16 and the kind::tf32 instantiations both assemble to it),",
         "UTMALDG/UTMASTG = TMA load/store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit -> mbarrier).  Template arguments of the conv kernel:",
         "`<BLOCK_N, NMAIN, 2SM, KIND (1 = f16, 0 = tf32), RING, SLOTS, HALO, WS>`.", "",
         "| kernel | instr | UTCHMMA (.2CTA) | other UTC*MMA | UTMALDG | UTMASTG | LDTM | UTCBAR | F2FP | SYNCS | BAR |", "|---|---|---|---|---|---|---|---|---|---|---|"]
tot = collections.Counter()
for k, c in C.items():
    name = re.sub(r"^void ", "", k).replace("dt::", "").replace("(dt::ConvParams)", "").replace("(int)", "").replace("(bool)", "")
    lines.append("| `%s` | %d | %d (%d) | %d | %d | %d | %d | %d | %d | %d | %d |" % (name[:96], c["instr"], c["UTCHMMA"], c["2CTA"], c["UTCother"], c["UTMALDG"], c["UTMASTG"],
                                                                                     c["LDTM"], c["UTCBAR"], c["F2FP"], c["SYNCS"], c["BAR"]))
    tot.update(c)
lines += ["", "Totals over the library (%d kernels): UTCHMMA %d (of which .2CTA %d), other UTC*MMA %d, UTMALDG %d, UTMASTG %d, LDTM %d, UTCBAR %d, F2FP %d." %
          (len(C), tot["UTCHMMA"], tot["2CTA"], tot["UTCother"], tot["UTMALDG"], tot["UTMASTG"], tot["LDTM"], tot["UTCBAR"], tot["F2FP"])]
ldd = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
deps = sorted(set(l.split()[0] for l in ldd.splitlines() if l.strip()))
lines += ["No cuBLAS / cuDNN / NCCL on the compute path: `ldd libdetectorch_b200.so` lists " + ", ".join("`%s`" % d for d in deps) + " (cudart is linked statically)."]
open(os.path.join(ROOT, "profiles", "r02_sass_summary.md"), "w").write("\n".join(lines) + "\n")
print(lines[-2]); print(lines[-1])
