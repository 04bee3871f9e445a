"""Summarise an ncu report: key raw metrics + top stall sites of the source page.  usage: ncu_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "gpu__dram_throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__warps_active.avg.pct", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu", "l1tex__lsu_writeback_active", "smsp__warp_issue_stalled"]
for i, h in enumerate(hdr):
    if any(h.startswith(w.strip()) for w in want):
        print(f"{h:75s}", [r[i] for r in rows[2:]][:3], rows[1][i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
tot = sum(int(r[si]) for r in body if len(r) > si and r[si].isdigit())
print("total samples", tot, "instructions", len(body))
agg = {}
for r in body:
    if len(r) <= si or not r[si].isdigit(): continue
    for i in range(len(h)):
        if h[i].startswith("stall_") and "Not Issued" not in h[i] and r[i] not in ("0", ""):
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:10])
top = sorted([(int(r[si]), k) for k, r in enumerate(body) if len(r) > si and r[si].isdigit()], reverse=True)[:n]
for s, k in top:
    r = body[k]
    reasons = {h[i][6:]: r[i] for i in range(len(h)) if h[i].startswith("stall_") and "Not Issued" not in h[i] and r[i] not in ("0", "")}
    print(f"{k:5d} {s:5d} {100*s/tot:5.1f}% x{r[ie]:>7s} {r[so][:60]:60s} {reasons}")
