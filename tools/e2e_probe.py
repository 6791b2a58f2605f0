"""Where the host-buffer entry point's time goes: gqe_score_grouped_host per call, compose on/off."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphqembed_b200 as gqe
from graphqembed_b200 import _lib
from graphqembed_b200.workloads import make_workload
import bench
device = torch.device("cuda", 0)
wl = make_workload(seed=0)
tables, rels, pre, post = bench.device_parameters(wl, torch, device, seed=1234)
lookup = gqe.RowLookup(wl.kg.node_ids)
mode_ids = {m: i for i, m in enumerate(wl.kg.modes)}
rel_ids = {r: i for i, r in enumerate(wl.kg.rel_keys)}
segs, anchor_rows, pair_rows = wl.lower(lookup, mode_ids, rel_ids)
ctx = gqe.Context(0, torch.cuda.current_stream().cuda_stream)
ctx.bind_tables([t.data_ptr() for t in tables], [t.size(0) for t in tables], wl.d)
ctx.bind_relations(0, [r.data_ptr() for r in rels], wl.d)
ctx.bind_intersection(0, [p.data_ptr() for p in pre], [p.data_ptr() for p in post], wl.d)
h_anchor = torch.from_numpy(anchor_rows).pin_memory(); h_pairs = torch.from_numpy(pair_rows).pin_memory(); h_loss = torch.zeros(1).pin_memory()
d_anchor, d_pairs, d_loss = h_anchor.to(device), h_pairs.to(device), torch.zeros(1, device=device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
nq = wl.n_queries
def host(): ctx._check(ctx._lib.gqe_score_grouped_host(ctx._h, segs, len(segs), nq, h_anchor.data_ptr(), h_pairs.data_ptr(), 2, None, 1.0, h_loss.data_ptr()))
def dev(): ctx.score_grouped_device(segs, nq, d_anchor.data_ptr(), d_pairs.data_ptr(), 2, None, 1.0, d_loss.data_ptr())
def timeit(fn, n=50, do_flush=True, sync_after=True):
    tot = 0.0
    for _ in range(n):
        if do_flush: flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn()
        if sync_after: torch.cuda.synchronize()
        tot += time.perf_counter() - t0
    return tot / n * 1e6
for mode in ("auto", "off"):
    ctx.set_compose(mode)
    for _ in range(5): host(); dev()
    print("compose=%-5s host-entry %.1f us | device-entry + sync %.1f us | device-entry CPU-only (no sync) %.1f us | host-entry warm L2 %.1f us" % (
        mode, timeit(host), timeit(dev), timeit(dev, sync_after=False), timeit(host, do_flush=False)))
def h2d():
    d_anchor.copy_(h_anchor, non_blocking=True); d_pairs.copy_(h_pairs, non_blocking=True)
print("H2D of the index arrays + sync: %.1f us" % timeit(h2d))
