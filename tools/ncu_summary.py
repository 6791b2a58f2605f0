"""Summarise an .ncu-rep (ncu --set full) as markdown for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [title] > profiles/rNN_x.md

Per profiled launch: grid / block / registers / duration, DRAM bytes (the bench's
`roofline.traffic`), pipe and memory utilisation, then for the first launch the
SASS evidence (tcgen05 / TMA mnemonics with their executed-instruction counts)
and the top stall sites.
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % (of active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
MARKS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCATOMSWS",
         "SYNCS", "HMMA", "LDG.E.128", "LDG.E.128.CONSTANT")


def run(args):
    return subprocess.run(["ncu", "-i"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    rows = list(csv.reader(run([rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# %s\n" % title)
    print("Source report: `%s` (ncu --set full --clock-control none --import-source on; per-launch values, "
          "cold-cache and serialised under the profiler -- not bench numbers).\n" % rep.split("/")[-1])
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        print("## launch %s: `%s`\n" % (r[ix["ID"]], r[ix["Kernel Name"]]))
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in ix and r[ix[k]] != "":
                print("| %s (`%s`) | %s %s |" % (label, k, r[ix[k]], units[ix[k]]))
        print()
    out = run([rep, "--page", "source", "--csv", "--print-source", "sass"])
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        return
    hdr = rows[starts[0] + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    end = starts[1] if len(starts) > 1 else len(rows)
    data = [r for r in rows[starts[0] + 2:end] if len(r) == len(hdr)]
    print("## SASS evidence, first launch (`%s`)\n" % rows[starts[0]][1])
    counts = {}
    for r in data:
        src = r[ix["Source"]].strip()
        parts = src.split()
        op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")
        for m in MARKS:
            if op.startswith(m):
                c = counts.setdefault(m, [0, 0])
                c[0] += 1
                c[1] += int(r[ix["Instructions Executed"]] or 0)
    print("| SASS mnemonic | static sites | warp-level executions |\n|---|---|---|")
    for m in MARKS:
        if m in counts:
            print("| `%s` | %d | %d |" % (m, counts[m][0], counts[m][1]))
    print()
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    agg = {}
    for r in data:
        for s in stalls:
            agg[s] = agg.get(s, 0) + int(r[ix[s]] or 0)
    print("Warp-state samples: %d total. By stall reason: %s\n" % (
        tot, ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1))
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8] if v)))
    print("Top stall sites:\n\n| samples | share | SASS | dominant stall |\n|---|---|---|---|")
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]
    for r in top:
        n = int(r[ix["# Samples"]] or 0)
        st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[0]
        print("| %d | %.1f%% | `%s` | %s |" % (n, 100.0 * n / max(tot, 1), r[ix["Source"]].strip()[:80], st[1]))


if __name__ == "__main__":
    main()
