"""Summarise an ncu report of gqe_fused_tc: headline metrics + stall samples by source line range."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum"]
for i, h in enumerate(hdr):
    if h in keys:
        print("%-70s %s %s" % (h, r[i], units[i]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], stdout=subprocess.PIPE,
                     text=True).stdout
rows = list(csv.reader(src.splitlines()))
# find header row with "# Samples"
hi = [i for i, x in enumerate(rows) if "# Samples" in x][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0
lines = []
for x in rows[hi + 1:]:
    if len(x) != len(hdr):
        continue
    try:
        n = int(x[ix["# Samples"]] or 0)
    except ValueError:
        continue
    tot += n
    lines.append((n, x[0], x[1]))
print("total samples", tot)
for n, ln, text in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%6d %5.1f%% L%-4s %s" % (n, 100.0 * n / max(tot, 1), ln, text.strip()[:110]))
