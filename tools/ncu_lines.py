"""Warp-state samples per CUDA source line (ncu --import-source on report), grouped by file.
    python tools/ncu_lines.py report.ncu-rep [min_samples]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
cur = None
hdr = None
tot = 0
rows = []
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        si = r.index("# Samples")
        stall_cols = [(i, h[6:]) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        n = int(r[si] or 0)
    except ValueError:
        continue
    tot += n
    st = sorted(((int(r[i] or 0), name) for i, name in stall_cols), reverse=True)[:2]
    rows.append((cur, int(r[0]), n, r[1].strip()[:100], " ".join("%s:%d" % (b, a) for a, b in st if a)))
print("total samples", tot)
for f, ln, n, text, st in rows:
    if n >= thr:
        print("%-12s L%-4d %6d %5.1f%%  %-100s %s" % (f, ln, n, 100.0 * n / max(tot, 1), text, st))
