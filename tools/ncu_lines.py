"""Top source lines by stall samples from an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict(); cur = None; fname = ""
hdr = None
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip()[:110]); agg.setdefault(cur, [0, 0, collections.Counter()])
    elif cur is not None:
        s = int(r[si]) if r[si].isdigit() else 0; n = int(r[ii]) if r[ii].isdigit() else 0
        a = agg[cur]; a[0] += s; a[1] += n
        if s: a[2][r[3].strip().split()[0] if not r[3].strip().startswith("@") else r[3].strip().split()[1]] += s
tot = sum(a[0] for a in agg.values())
print(f"total samples {tot}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    ops = ", ".join(f"{o}:{c}" for o, c in a[2].most_common(4))
    print(f"{100*a[0]/tot:5.1f}% {a[1]:9d} inst  {k[0]}:{k[1]:<4d} {k[2]}\n          [{ops}]")
