import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tot = collections.Counter(); smp = collections.Counter(); thr = collections.Counter()
cur = hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = {}; [hdr.setdefault(h, k) for k, h in enumerate(r)]; continue
    if hdr is None or not r[0].isdigit(): continue
    g = lambda c: int(r[hdr[c]]) if c in hdr and r[hdr[c]].isdigit() else 0
    tot[(cur, int(r[0]))] += g("Instructions Executed"); smp[(cur, int(r[0]))] += g("# Samples"); thr[(cur, int(r[0]))] += g("Thread Instructions Executed")
T = sum(tot.values()); S = sum(smp.values())
print("total", T, S, "thread/warp", sum(thr.values()) / max(T, 1))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
for (f, l), v in sorted(tot.items()):
    if v and lo <= l <= hi: print(f, l, f"{100*v/T:5.2f}% inst  {100*smp[(f,l)]/S:5.2f}% smp  lanes {thr[(f,l)]/max(v,1):4.1f}")
