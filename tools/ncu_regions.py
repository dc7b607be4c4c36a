"""Dynamic instruction / stall-sample shares per source region of nmf_step.cuh from an .ncu-rep (cuda,sass source view).
Regions are found by the function / section markers in the source, so the table survives edits."""
import collections, csv, io, re, subprocess, sys
from pathlib import Path
rep = sys.argv[1]
src = Path(__file__).resolve().parent.parent / "flygym_b200/csrc/nmf_step.cuh"
marks = []   # (line, name)
for n, l in enumerate(src.read_text().split("\n"), 1):
    m = re.match(r"__device__ __forceinline__ [\w ]*?(\w+)\(", l) or re.match(r"template <.*> __device__ __forceinline__ \w+ (\w+)\(", l)
    if m: marks.append((n, m.group(1)))
    m = re.match(r"\s*// (A|B|C)\. ", l)
    if m: marks.append((n - 1, "step:" + m.group(1)))
    if "---- exact line search" in l: marks.append((n, "step:linesearch"))
    if "---- optional outputs" in l: marks.append((n, "step:outputs"))
    if "---- advance:" in l: marks.append((n, "step:advance"))
marks.sort()
def region(ln):
    name = "?"
    for a, nm in marks:
        if a <= ln: name = nm
        else: break
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tot, smp, noi, sb = (collections.Counter() for _ in range(4))
cur = hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = {}; [hdr.setdefault(h, k) for k, h in enumerate(r)]; continue
    if hdr is None or not r[0].isdigit(): continue
    g = lambda c: int(r[hdr[c]]) if r[hdr[c]].isdigit() else 0
    key = region(int(r[0])) if cur == "nmf_step.cuh" else cur
    tot[key] += g("Instructions Executed"); smp[key] += g("# Samples"); noi[key] += g("stall_no_inst"); sb[key] += g("stall_short_sb")
T, S, N = sum(tot.values()), sum(smp.values()), sum(noi.values())
print(f"total warp instructions {T}, samples {S}, no_inst samples {N}")
for k, v in tot.most_common(32):
    print(f"{k:26s} inst {100*v/T:5.1f}%  samples {100*smp[k]/S:5.1f}%  no_inst {100*noi[k]/max(1,N):5.1f}%  short_sb {100*sb[k]/max(1,sum(sb.values())):5.1f}%")
