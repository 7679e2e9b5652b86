"""Opcode evidence that the shipped library is Blackwell-native: per kernel of
czech-contriever_b200/lib/libb2ip.so, the number of tcgen05 MMA (UTCHMMA / UTCQMMA ...), TMEM
load (LDTM), TMA (UTMALDG / UTMASTG), tensor-core barrier (UTCBAR) and legacy mma.sync (HMMA)
instructions in the sm_100a SASS (`cuobjdump -sass`).  CPU only.

    python tools/sass_summary.py > profiles/r2_sass_opcodes.md
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "czech-contriever_b200", "lib", "libb2ip.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "IMMA", "DFMA", "F2F.F64.F32", "ATOM", "RED"]
arch = re.findall(r"arch = (sm_\w+)", sass)
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in pats:
            if op.startswith(p):
                counts[cur][p] += 1
                if p == "UTCHMMA" and ".2CTA" in op:
                    counts[cur]["UTCHMMA.2CTA"] += 1
                if p == "UTMALDG" and ".2CTA" in op:
                    counts[cur]["UTMALDG.2CTA"] += 1
print("# SASS opcode summary of libb2ip.so (final code of round 2)\n")
print(f"`cuobjdump -sass czech-contriever_b200/lib/libb2ip.so`, architectures in the file: {sorted(set(arch))}.")
print("UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit,")
print("SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0).\n")
cols = ["_total", "UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTCBAR", "SYNCS", "HMMA", "DFMA", "F2F.F64.F32", "ATOM", "RED"]
print("| kernel | " + " | ".join(c.replace("_total", "instructions") for c in cols) + " |")
print("|---|" + "---:|" * len(cols))
for k, c in counts.items():
    print(f"| `{k[:90]}` | " + " | ".join(str(c.get(col, 0)) for col in cols) + " |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"\nTotals: UTCHMMA {tot['UTCHMMA']} (of which .2CTA {tot['UTCHMMA.2CTA']}), LDTM {tot['LDTM']}, UTMALDG {tot['UTMALDG']} "
      f"(.2CTA {tot['UTMALDG.2CTA']}), UTCBAR {tot['UTCBAR']}, HMMA {tot['HMMA']}.")
