"""A/B timing of the fused kernel on the headline mix: rows vs node ids, weight cache on (run under gpurun)."""
import os
import sys

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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=device)     # 256 MiB read after the write: evicts the dirty lines
FLUSH = os.environ.get("AB_FLUSH", "write")                                # write | wr (write, then read) | none


def do_flush():
    if FLUSH != "none":
        flush.zero_()
    if FLUSH == "wr":
        flush_rd.sum()


d_loss = torch.zeros(1, device=device)
stream = torch.cuda.current_stream()
for nodes in (False, True, False, True):
    segs, a, t = wl.node_arrays(mode_ids, rel_ids) if nodes else wl.lower(lookup, mode_ids, rel_ids)
    ctx = gqe.Context(0, stream.cuda_stream)
    ctx.bind_tables([x.data_ptr() for x in tables], [x.size(0) for x in tables], wl.d)
    ctx.bind_relations(_lib.DECODER_ID[wl.decoder], [r.data_ptr() for r in rels], wl.d)
    if wl.inter.endswith("-simple"):
        ctx.bind_intersection(_lib.INTER_ID[wl.inter], None, None, wl.d)
    else:
        ctx.bind_intersection(_lib.INTER_ID[wl.inter], [p.data_ptr() for p in pre], [p.data_ptr() for p in post], wl.d)
    maps = lookup.device_maps(wl.kg.modes, [x.size(0) for x in tables], device)
    ctx.bind_node_maps(maps[0], maps[1], maps[2])
    da, dt = torch.from_numpy(a).to(device), torch.from_numpy(t).to(device)

    def step():
        ctx.score_grouped_device(segs, wl.n_queries, da.data_ptr(), dt.data_ptr(), 2, None, 1.0, d_loss.data_ptr(), nodes=nodes)
    for _ in range(5):
        do_flush(); step()
    torch.cuda.synchronize()
    n = 100
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in evs:
        do_flush(); e0.record(stream); step(); e1.record(stream)
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    print("flush=%s " % FLUSH, end="")
    print("%s %s: mean %.4f ms  median %.4f  min %.4f  loss %.7f" % (name, "nodes" if nodes else "rows ", sum(ts) / n, ts[n // 2], ts[0], d_loss.item()))
