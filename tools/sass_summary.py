"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG, stmatrix -> STSM, legacy mma.sync -> HMMA) in the shipped library.
  python tools/sass_summary.py [path/to/lib.so] > profiles/r02/sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "ddiffpg_b200/libddiffpg_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "STSM", "LDSM", "MUFU", "RED", "ATOM", "SYNCS", "USETMAXREG")
counts, total = collections.defaultdict(collections.Counter), collections.Counter()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = fn.replace("(anonymous namespace)::", "")
        fn = re.sub(r"^void ", "", re.sub(r"\(.*", "", fn)).split("::")[-1] or m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for w in want:
            if op.startswith(w):
                counts[fn][w] += 1
print(f"SASS mnemonic counts per kernel of {lib} (cuobjdump -sass; arch sm_100a)")
print(f"{'kernel':44s} {'instr':>7s} " + " ".join(f"{w:>8s}" for w in want))
for fn in sorted(total, key=lambda k: -total[k]):
    c = counts[fn]
    if not any(c[w] for w in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA")):
        continue
    print(f"{fn[:44]:44s} {total[fn]:7d} " + " ".join(f"{c[w]:8d}" for w in want))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'all kernels':44s} {sum(total.values()):7d} " + " ".join(f"{tot[w]:8d}" for w in want))
