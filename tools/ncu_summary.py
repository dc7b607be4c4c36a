"""Summarise an .ncu-rep (raw + source pages) into a small text report for profiles/."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
print("== raw metrics ==")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); stall = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ti = ts = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]].strip())
    op = (m.group(2) if m else "?").split(".")[0]
    n = int(r[idx["Instructions Executed"]] or 0); s = int(r[idx["# Samples"]] or 0)
    ops[op] += n; samples[op] += s; ti += n; ts += s
    for c in stall_cols: stall[c] += int(r[idx[c]] or 0)
print(f"== opcode mix: total warp instructions {ti}, samples {ts} ==")
for op, n in ops.most_common(22):
    print(f"{op:14s} inst {n:11d} {100*n/ti:5.1f}%   samples {100*samples[op]/max(ts,1):5.1f}%")
print("== stall reasons (samples) ==")
for k, v in sorted(stall.items(), key=lambda x: -x[1])[:10]:
    print(f"{k:28s} {v:8d} {100*v/max(ts,1):5.1f}%")
