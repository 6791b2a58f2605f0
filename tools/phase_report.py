"""Where a tile's time goes inside gqe_fused_tc: decode the phase stamps of
gqe_debug_set_phase_log for one step of a bench workload (run under gpurun).

    python tools/phase_report.py [workload] [formulas_per_structure]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphqembed_b200 as gqe  # noqa: E402
from graphqembed_b200 import _lib  # noqa: E402
from graphqembed_b200.workloads import DEFAULT_WORKLOAD, make_workload  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_WORKLOAD
fps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
device = torch.device("cuda", 0)
wl = make_workload(name, seed=0, formulas_per_structure=fps)
tables, rels, pre, post = bench.device_parameters(wl, torch, device, seed=1234)
lookup = gqe.RowLookup(wl.kg.node_ids)
mode_ids = {m: i for i, m in enumerate(wl.kg.modes)}
rel_ids = {r: i for i, r in enumerate(wl.kg.rel_keys)}
NODES = os.environ.get("GQE_NODES", "0") == "1"      # node ids mapped inside the kernel instead of host-lowered rows
if NODES:
    segs, anchor_rows, pair_rows = wl.node_arrays(mode_ids, rel_ids)
else:
    segs, anchor_rows, pair_rows = wl.lower(lookup, mode_ids, rel_ids)
ctx = gqe.Context(0, torch.cuda.current_stream().cuda_stream)
ctx.bind_tables([t.data_ptr() for t in tables], [t.size(0) for t in tables], wl.d)
if NODES:
    _maps = lookup.device_maps(wl.kg.modes, [t.size(0) for t in tables], device)
    ctx.bind_node_maps(_maps[0], _maps[1], _maps[2])
ctx.bind_relations(_lib.DECODER_ID[wl.decoder], [r.data_ptr() for r in rels], wl.d)
ctx.bind_intersection(_lib.INTER_ID[wl.inter], [p.data_ptr() for p in pre], [p.data_ptr() for p in post], wl.d)
d_anchor, d_pairs = torch.from_numpy(anchor_rows).to(device), torch.from_numpy(pair_rows).to(device)
d_loss = torch.zeros(1, device=device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
n_tiles = 4096
log = torch.zeros(n_tiles * 32, dtype=torch.int64, device=device)


def step():
    ctx.score_grouped_device(segs, wl.n_queries, d_anchor.data_ptr(), d_pairs.data_ptr(), 2, None, 1.0, d_loss.data_ptr(),
                             nodes=NODES)


for _ in range(3):
    flush.zero_()
    step()
torch.cuda.synchronize()
ctx.debug_set_phase_log(log.data_ptr(), n_tiles)
flush.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
ctx.debug_set_phase_log(None, 0)
print("step %.1f us" % (e0.elapsed_time(e1) * 1e3))
raw = log.cpu().numpy().astype(np.uint64).reshape(n_tiles, 32)
# CTA-level stamps (last 256 records, record n_tiles-1-cta): globaltimer ns / clock64 pairs
cta = raw[::-1][:256].astype(np.int64)
cta = cta[cta[:, 0] != 0]
if len(cta):
    gt = cta[:, 0:10:2].astype(np.float64)
    t0 = gt[:, 0].min()
    names = ["entry", "set-up done", "first tile taken", "tile loop left", "exit"]
    print("CTA stamps (%d CTAs), us after the first CTA's entry (globaltimer): " % len(cta))
    for i, nm in enumerate(names):
        col = (gt[:, i] - t0) / 1e3
        print("  %-17s min %7.1f  avg %7.1f  max %7.1f" % (nm, col.min(), col.mean(), col.max()))
    ck = cta[:, 1:10:2].astype(np.float64)
    mhz = (ck[:, 3] - ck[:, 1]) / np.maximum(gt[:, 3] - gt[:, 1], 1.0) * 1e3
    print("  SM clock inside the tile loop (clock64 / globaltimer): median %.0f MHz (tile times below assume %.0f)" % (np.median(mhz), 1965.0))
    print("  kernel span (first entry -> last exit) %.1f us; event-timed step above includes pack/compose" % ((gt[:, 4].max() - t0) / 1e3))
raw = raw.copy()
raw[-256:] = 0
tags = (raw >> np.uint64(56)).astype(np.int64)
clk = (raw & np.uint64((1 << 56) - 1)).astype(np.int64)
MHZ = 1965.0
NAMES = {2: "gather", 3: "wait MMA", 4: "epilogue", 5: "acc->smem transpose", 6: "score", 7: "loss reduce", 8: "deferred score of the previous tile", 9: "hand-over + prefetch issue"}
STRUCT = ["1-chain", "2-chain", "3-chain", "2-inter", "3-inter", "3-inter_chain", "3-chain_inter"]
by_struct = {}
for t in range(n_tiles):
    if tags[t, 0] == 0:
        continue
    s = (tags[t, 0] - 1) // 16
    acc = by_struct.setdefault(s, {"n": 0, "total": 0.0, "first_gather": 0.0})
    acc["n"] += 1
    prev = clk[t, 0]
    first = True
    for i in range(1, 31):
        if tags[t, i] == 0:
            break
        dt = (clk[t, i] - prev) / MHZ
        k = int(tags[t, i])
        if k == 2 and first:
            acc["first_gather"] += dt
            first = False
        else:
            acc[k] = acc.get(k, 0.0) + dt
        acc["cnt%d" % k] = acc.get("cnt%d" % k, 0) + 1
        prev = clk[t, i]
    acc["total"] += (prev - clk[t, 0]) / MHZ
grand = sum(v["total"] for v in by_struct.values())
print("tiles: %d, sum of tile times %.0f us -> %.1f us per SM over 148 SMs" % (sum(v["n"] for v in by_struct.values()), grand, grand / 148))
for s in sorted(by_struct):
    v = by_struct[s]
    n = v["n"]
    print("%-14s tiles=%4d  avg tile %.1f us:" % (STRUCT[s], n, v["total"] / n), end="")
    print("  first gather %.1f" % (v["first_gather"] / n), end="")
    for k in (2, 9, 8, 3, 4, 5, 6, 7):
        if k in v:
            print(" | %s %.1f (x%.1f)" % (NAMES[k], v[k] / n, v.get("cnt%d" % k, 0) / n - (1 if k == 2 else 0)), end="")
    print()
tot = {}
for v in by_struct.values():
    tot["first gather"] = tot.get("first gather", 0) + v["first_gather"]
    for k in NAMES:
        if k in v:
            tot[NAMES[k]] = tot.get(NAMES[k], 0) + v[k]
print("share of summed tile time:", ", ".join("%s %.1f%%" % (k, 100 * x / grand) for k, x in tot.items()))

# per-SM view: tiles grouped by the SM that ran them (slot 31 = %smid)
smid = raw[:, 31].astype(np.int64)
valid = tags[:, 0] != 0
span, busy = [], []
for sm in np.unique(smid[valid]):
    idx = np.flatnonzero(valid & (smid == sm))
    first = clk[idx, 0].min()
    last = max(clk[t][(tags[t, :31] != 0).sum() - 1] for t in idx)
    span.append((last - first) / MHZ)
    busy.append(sum((clk[t][(tags[t, :31] != 0).sum() - 1] - clk[t, 0]) for t in idx) / MHZ)
print("per SM (%d SMs): first-stamp-to-last-stamp span avg %.1f max %.1f min %.1f us; in-tile time avg %.1f us; tiles/SM avg %.2f" % (
    len(span), np.mean(span), np.max(span), np.min(span), np.mean(busy), valid.sum() / len(span)))
