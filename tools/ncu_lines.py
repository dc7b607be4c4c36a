"""Top source lines of one file by executed warp instructions / stall samples from an .ncu-rep (needs -lineinfo)."""
import collections, csv, io, subprocess, sys
from pathlib import Path
rep, fname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
inst, smp = collections.Counter(), collections.Counter()
files = collections.Counter()
cur = hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = {}; [hdr.setdefault(h, k) for k, h in enumerate(r)]; continue
    if hdr is None or not r[0].isdigit(): continue
    g = lambda c: int(r[hdr[c]]) if c in hdr and r[hdr[c]].isdigit() else 0
    files[cur] += g("Instructions Executed")
    if cur == fname: inst[int(r[0])] += g("Instructions Executed"); smp[int(r[0])] += g("# Samples")
T, S = sum(files.values()), sum(smp.values())
print("instructions by file:", {k: f"{100*v/T:.1f}%" for k, v in files.most_common(6)}, "total", T)
src = (Path(__file__).resolve().parent.parent / "flygym_b200/csrc" / fname).read_text().split("\n")
order = smp.most_common(top) if len(sys.argv) > 4 and sys.argv[4] == "samples" else inst.most_common(top)
for ln, _ in order:
    v = inst[ln]
    print(f"{ln:5d} inst {100*v/T:5.1f}% smp {100*smp[ln]/max(1,S):5.1f}%  {src[ln-1].strip()[:110]}")
