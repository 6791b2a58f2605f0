"""Stall samples per SASS instruction of the first kernel in an ncu report (markers + hot spots)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
end = starts[1] if len(starts) > 1 else len(rows)
data = [r for r in rows[2:end] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
cum = 0
for r in data:
    n = int(r[ix["# Samples"]] or 0)
    cum += n
    src = r[ix["Source"]]
    if n >= thr or any(k in src for k in ("SYNCS", "UTC", "UBLKCP", "BAR.", "LDTM", "STTM", "EXIT", "UBLKPF", "CCTL")):
        st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print("%6d cum%5.1f%% %-6s %-90s %s" % (n, 100.0 * cum / tot, r[ix["Address"]][-5:], src[:90],
                                               " ".join("%s:%d" % (b, a) for a, b in st if a)))
