import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []
cur = hdr = None
agg = {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = {}; [hdr.setdefault(h, k) for k, h in enumerate(r)]; stall_cols = [h for h in r if h.startswith("stall_")]; continue
    if hdr is None or not r[0].isdigit(): continue
    g = lambda c: int(r[hdr[c]]) if c in hdr and r[hdr[c]].isdigit() else 0
    key = (cur, int(r[0]))
    a = agg.setdefault(key, collections.Counter())
    a["inst"] += g("Instructions Executed"); a["smp"] += g("# Samples")
    for c in stall_cols: a[c] += g(c)
S = sum(a["smp"] for a in agg.values()); T = sum(a["inst"] for a in agg.values())
top = sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for (f, l), a in top:
    st = sorted(((v, k) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:3]
    print(f"{f}:{l}  smp {100*a['smp']/S:5.2f}%  inst {100*a['inst']/T:5.2f}%  " + "  ".join(f"{k[6:]} {100*v/max(a['smp'],1):.0f}%" for v, k in st))
