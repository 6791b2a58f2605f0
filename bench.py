#!/usr/bin/env python
"""bench.py -- conjunctive queries scored/sec on synthetic Bio-shaped batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch: every query of the batch
is scored against its positive and one negative target and folded into the
margin loss (what reference model.py:112-127 computes per batch).  At N=1 the
headline workload is BASELINE.json configs[3] (Bio KG full mix, d=256,
batch=65536) -- the metric string names no single config, so the largest
single-GPU configuration is used; every other config rides in `configs`.

`value`    whole-job throughput with the NODE-ID arrays already resident in HBM
           (gqe_score_grouped_nodes_device: node id -> row lookup, gather, ...,
           loss, one kernel), CUDA events on the launching stream, L2 flushed
           before every step, max over ranks.  The operator matrices did not change
           between steps, so their packed images are cached (steady state of
           eval / inference); `weights_live` is the same step with the cache off.
`e2e`      same metric through gqe_score_grouped_nodes_host: pinned HOST node-id
           buffers in, HOST loss out, copies + stream sync inside the timed call --
           the span the reference arm pays (index building included).
           `e2e.from_store` / `e2e.from_query_objects` go through the Python
           operator surface (QueryEncoderDecoder.margin_loss) from a QueryStore
           slice / from lists of Query objects, negative sampling included.
`roofline` algorithmic flops (SURVEY.md 8d, x3 for the split-bf16 scheme) / the
           fused kernel's own duration vs the MEASURED burst bf16 peak; the flops
           the tensor pipe really issued after pre-composition beside it.
`configs`  BASELINE.json configs[0..4] (+ min aggregator, + 66-formula mix), each
           with device / e2e time and both roofline views.
`hbm_bound` TransE / DistMult + SimpleSetIntersection on the 10 M-node KG: the
           HBM-gather-bound operators against the measured copy bandwidth.
`cpu_baseline` the oracle port of the reference path timed on this box's cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "conjunctive queries scored/sec"
UNIT = "queries/s"
HBM_FALLBACK_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback
BF16_FALLBACK_TFLOPS = 1590.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("native", "reference"), default="native")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--formulas-per-structure", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=12288, help="queries in the CPU-baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tables", choices=("replicated", "p2p", "staged"), default=None,
                    help="table placement (default: replicated; p2p for the 10M-node workload at N>1)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded 10M-node leg at N>1")
    ap.add_argument("--no-eval-shape", action="store_true", help="skip the 1 query x 101 targets leg at N=1")
    ap.add_argument("--no-train-step", action="store_true", help="skip the configs[0] training-step leg at N=1")
    ap.add_argument("--with-sharded", action="store_true", help="(kept for compatibility; the 10M-node table now runs at N=1 by default)")
    ap.add_argument("--sharded-formulas", type=int, default=4, help="formulas per structure of the 10M-node leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (configs, hbm_bound) at N=1")
    ap.add_argument("--extra-steps", type=int, default=40, help="timed steps of each secondary workload")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        # the fused kernel is a ~0.1 ms launch timed alone behind an L2 flush: the BURST figure applies
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": HBM_FALLBACK_GBS, "bf16_tflops": BF16_FALLBACK_TFLOPS, "bf16_tflops_sustained": BF16_FALLBACK_TFLOPS,
            "source": "fallback"}


# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons via NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.active, self._stop_evt = index, [], False, threading.Event()
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES remapping when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self.active, mhz, mask))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        under = [s for s in self.samples if s[0]] or self.samples
        mhz = sorted(s[1] for s in under)
        mask = 0
        for s in under:
            mask |= s[2]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(under)}


# ---------------------------------------------------------------------------
def device_parameters(wl, torch, device, seed, owner=None, rank=0):
    """Random-init parameters of the reference's architecture, on the device:
    tables N(0, 1/d) (bio/data_utils.py:17-19), xavier-uniform relation / pre /
    post matrices (decoders.py:139,282,285).  Every tensor has its own seed, so
    a rank that holds only its shard (``owner[m] == rank``) holds exactly the
    values the single-GPU run holds for that mode."""
    import math
    d, kg = wl.d, wl.kg

    def gen(k):
        return torch.Generator(device=device).manual_seed(seed * 1000003 + k)

    tables = []
    for i, m in enumerate(kg.modes):
        if owner is not None and owner[i] != rank:
            tables.append(None)
        else:
            tables.append(torch.randn(kg.sizes[m] + 2, d, generator=gen(i), device=device) * (1.0 / d))
    bound = math.sqrt(6.0 / (2 * d))
    uni = lambda k: (torch.rand(d, d, generator=gen(k), device=device) * 2 - 1) * bound
    if wl.decoder == "bilinear":
        rels = [uni(1000 + i) for i in range(len(kg.rel_keys))]
    else:      # relation vectors U(+-6/sqrt(d)) (decoders.py:195,224)
        rels = [(torch.rand(d, generator=gen(1000 + i), device=device) * 2 - 1) * (6.0 / math.sqrt(d))
                for i in range(len(kg.rel_keys))]
    pre = [uni(2000 + i) for i in range(len(kg.modes))]
    post = [uni(3000 + i) for i in range(len(kg.modes))]
    return tables, rels, pre, post


class Timed(object):
    """K timed steps after W warm-ups: CUDA events on the launching stream, L2
    flushed before every step, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, torch, dist, device, stream, flush):
        self.torch, self.dist, self.device, self.stream, self.flush = torch, dist, device, stream, flush

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.device)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def device_ms(self, step, steps, warmup, sampler=None):
        torch = self.torch
        for _ in range(warmup):
            self.flush.zero_()
            step()
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        if sampler is not None:
            sampler.active = True
        for e0, e1 in evs:
            self.flush.zero_()
            e0.record(self.stream)
            step()
            e1.record(self.stream)
        self.barrier()
        if sampler is not None:
            sampler.active = False
        return self.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs))[0] / steps

    def host_ms(self, step, steps, warmup=3):
        for _ in range(warmup):
            step()
        self.barrier()
        tot = 0.0
        for _ in range(steps):
            self.flush.zero_()
            self.torch.cuda.synchronize(self.device)
            t0 = time.perf_counter()
            step()
            tot += time.perf_counter() - t0
        self.barrier()
        return self.max_over_ranks(tot * 1e3)[0] / steps


def measure(args, name, tables_mode, tm, rank, world, local_rank, sampler=None, formulas_per_structure=1, steps=None,
            parity_slice=0, want_live=True, routed=False):
    """One workload on this process group.  tables_mode:
         replicated  every rank holds every table; pure data parallel, no data-path collective
         p2p         tables sharded by node type, peers mapped with CUDA IPC, the fused kernel
                     reads remote rows over NVLink itself (a TMA helper warp stages them a tile ahead; GQE_STAGE=0: in place)
         staged      tables sharded by node type, rows fetched by an NCCL all-to-all exchange
                     into staging tables before the same fused kernel runs
       The index arrays are NODE IDS (int32) -- mapped to table rows inside the kernels through the
       bound node maps -- except on the staged path, whose indices are positions in the exchange
       buffers.  parity_slice > 0: score that many queries of every formula a second time against
       the full, unsharded tables held on this one GPU and require bit-equal scores."""
    import numpy as np
    import torch

    import graphqembed_b200 as gqe
    from graphqembed_b200 import _lib, sharded
    from graphqembed_b200.workloads import make_workload

    steps = steps or args.steps
    device, stream, dist = tm.device, tm.stream, tm.dist
    from graphqembed_b200.workloads import _kg_cached
    kg = _kg_cached("large" if name.startswith("synth-10m") else "bio")
    n_modes = len(kg.modes)
    owner = None if tables_mode == "replicated" else sharded.owner_by_node_type(n_modes, world)
    # routed: this rank scores the formulas whose TARGET node type it owns (sharded.route_by_target_mode)
    tmodes = [kg.modes[m] for m in sharded.modes_owned_by(owner, rank)] if (routed and owner is not None) else None
    wl = make_workload(name, seed=rank, kg=kg, formulas_per_structure=formulas_per_structure, target_modes=tmodes)
    tables, rels, pre, post = device_parameters(wl, torch, device, seed=1234, owner=owner, rank=rank)
    lookup = gqe.RowLookup(kg.node_ids)
    mode_ids = {m: i for i, m in enumerate(kg.modes)}
    rel_ids = {r: i for i, r in enumerate(kg.rel_keys)}
    segs, anchor_nodes, pair_nodes = wl.node_arrays(mode_ids, rel_ids)
    nq = wl.n_queries
    rows = [kg.sizes[m] + 2 for m in kg.modes]
    deepsets = not wl.inter.endswith("-simple")

    def new_context(ptrs, nrows, node_maps=True):
        c = gqe.Context(local_rank, stream.cuda_stream)
        c.bind_relations(_lib.DECODER_ID[wl.decoder], [r.data_ptr() for r in rels], wl.d)
        if deepsets:
            c.bind_intersection(_lib.INTER_ID[wl.inter], [p.data_ptr() for p in pre], [p.data_ptr() for p in post], wl.d)
        else:
            c.bind_intersection(_lib.INTER_ID[wl.inter], None, None, wl.d)
        c.bind_tables(ptrs, nrows, wl.d)
        if node_maps:
            maps = lookup.device_maps(kg.modes, nrows, device)
            c.bind_node_maps(maps[0], maps[1], maps[2])
        return c

    own_ptrs = [0 if t is None else t.data_ptr() for t in tables]
    own_rows = [0 if t is None else r for t, r in zip(tables, rows)]
    h_anchor = torch.from_numpy(anchor_nodes).pin_memory()
    h_pairs = torch.from_numpy(pair_nodes).pin_memory()
    h_loss = torch.zeros(1).pin_memory()
    d_loss = torch.zeros(1, device=device)
    remote_rows = 0
    peers = None
    d_scores = torch.empty((nq, 2), device=device) if parity_slice else None
    sub = None
    if parity_slice:
        sub = _lib.make_segments([(segs[i].plan, segs[i].query_begin,
                                   min(segs[i].query_end, segs[i].query_begin + parity_slice)) for i in range(len(segs))])

    if tables_mode in ("replicated", "p2p"):
        if tables_mode == "p2p" and world > 1:
            octx = new_context(own_ptrs, own_rows, node_maps=False)
            peers = sharded.PeerTables(octx, owner, rows, {m: t for m, t in enumerate(tables) if t is not None})
            ctx = new_context(peers.pointers(), rows)
            for c_mode, c_n, _, _, _ in sharded.chunks_of_segments(segs, nq, 2):
                if owner[c_mode] != rank:
                    remote_rows += c_n
        else:
            ctx = new_context(own_ptrs, own_rows)
        d_anchor, d_pairs = h_anchor.to(device), h_pairs.to(device)

        def step_device():
            ctx.score_grouped_device(segs, nq, d_anchor.data_ptr(), d_pairs.data_ptr(), 2, None, 1.0, d_loss.data_ptr(),
                                     nodes=True)

        def step_host():
            ctx._check(ctx._lib.gqe_score_grouped_nodes_host(ctx._h, segs, len(segs), nq, h_anchor.data_ptr(),
                                                             h_pairs.data_ptr(), 2, None, 1.0, h_loss.data_ptr()))

        def slice_scores():
            ctx.score_grouped_device(sub, nq, d_anchor.data_ptr(), d_pairs.data_ptr(), 2, d_scores.data_ptr(), 1.0, None,
                                     nodes=True)
        # bytes the host entry point really copies: per anchor slot only the query range of the
        # segments that use the slot (chains fill one of the three slots), plus the target pairs
        h2d = int(pair_nodes.nbytes)
        for k in range(_lib.GQE_MAX_ANCHORS):
            used = [(int(segs[i].query_begin), int(segs[i].query_end)) for i in range(len(segs))
                    if segs[i].plan.anchor_mode[k] >= 0]
            if used:
                h2d += 4 * (max(e for _, e in used) - min(b for b, _ in used))
        call = "gqe_score_grouped_nodes_host (pinned int32 NODE IDS in, fp32 loss out; lookup + scoring in one kernel)"
        launch_ctxs = (ctx,)
    else:
        # staged exchange: requests are table rows (the owner gathers them), lowered on the host
        _, anchor_rows, pair_rows = wl.lower(lookup, mode_ids, rel_ids)
        info = sharded.chunks_of_segments(segs, nq, 2)
        plan = sharded.ExchangePlan(owner, world, rank, [(m, n) for m, n, _, _, _ in info])
        ctx = new_context(own_ptrs, own_rows, node_maps=False)          # the owner-side gather reads these
        ex = sharded.RowExchange(plan, wl.d, sharded.device_gather(ctx), device=device)
        req, s_anchor, s_pairs = sharded.stage_grouped(plan, info, segs, anchor_rows, pair_rows, 2)
        h_req = torch.from_numpy(req).pin_memory()
        h_sanchor = torch.from_numpy(s_anchor).pin_memory()
        h_spairs = torch.from_numpy(s_pairs).pin_memory()
        d_req, d_sanchor, d_spairs = h_req.to(device), h_sanchor.to(device), h_spairs.to(device)
        e_req, e_sanchor, e_spairs = torch.empty_like(d_req), torch.empty_like(d_sanchor), torch.empty_like(d_spairs)
        for c_mode, c_n, _, _, _ in info:
            if owner[c_mode] != rank:
                remote_rows += c_n
        # a second context scores against the staging tables (the first keeps the shards bound)
        st_ptrs, st_rows = ex.staging_tables()
        sctx = new_context(st_ptrs, st_rows, node_maps=False)

        def step_device():
            ex.run(d_req)
            sctx.score_grouped_device(segs, nq, d_sanchor.data_ptr(), d_spairs.data_ptr(), 2, None, 1.0,
                                      d_loss.data_ptr())

        def step_host():
            e_req.copy_(h_req, non_blocking=True)
            e_sanchor.copy_(h_sanchor, non_blocking=True)
            e_spairs.copy_(h_spairs, non_blocking=True)
            ex.run(e_req)
            sctx.score_grouped_device(segs, nq, e_sanchor.data_ptr(), e_spairs.data_ptr(), 2, None, 1.0,
                                      d_loss.data_ptr())
            h_loss.copy_(d_loss, non_blocking=True)
            stream.synchronize()

        def slice_scores():
            ex.run(d_req)
            sctx.score_grouped_device(sub, nq, d_sanchor.data_ptr(), d_spairs.data_ptr(), 2, d_scores.data_ptr(), 1.0, None)
        h2d = int(req.nbytes + s_anchor.nbytes + s_pairs.nbytes)
        call = "request H2D + NCCL all-to-all exchange + gqe_score_grouped_device + loss D2H"
        launch_ctxs = (ctx, sctx)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        tm.flush.zero_()
        step_device()
    tm.barrier()
    loss_ref = float(d_loss.item())
    launches0 = sum(c.launch_count() for c in launch_ctxs)
    dev_ms = tm.device_ms(step_device, steps, 0, sampler)
    launches = (sum(c.launch_count() for c in launch_ctxs) - launches0) / float(steps)
    assert float(d_loss.item()) == loss_ref, "non-deterministic loss"
    if sampler is not None:
        sampler.stop()      # the NVML polling thread must not compete with the host-timed loop
    e2e_ms = tm.host_ms(step_host, steps)
    assert float(h_loss[0]) == loss_ref, "host entry point disagrees with the device one"
    tm.barrier()
    # the same step with the operator matrices treated as live (re-composed and re-packed every call:
    # what a training loop that updates them every step pays)
    live_ms = live_launches = None
    if want_live and wl.decoder == "bilinear" and tables_mode != "staged":
        ctx.set_weight_cache(False)
        l0 = ctx.launch_count()
        live_ms = tm.device_ms(step_device, max(10, steps // 2), 3)
        live_launches = (ctx.launch_count() - l0) / float(max(10, steps // 2) + 3)
        assert float(d_loss.item()) == loss_ref
        ctx.set_weight_cache(True)
    fused_ms = None
    if wl.decoder == "bilinear" and tables_mode != "staged" and wl.d in (128, 256):
        fused_ms = fused_kernel_span_ms(torch, ctx, step_device, tm, device)

    parity = None
    if parity_slice:
        # this rank's queries against the FULL tables held on this one GPU: must be bit-equal
        slice_scores()
        torch.cuda.synchronize(device)
        got = d_scores.clone()
        full, _, _, _ = device_parameters(wl, torch, device, seed=1234)
        fctx = new_context([t.data_ptr() for t in full], rows)
        d_a2, d_p2 = h_anchor.to(device), h_pairs.to(device)
        want = torch.empty_like(d_scores)
        fctx.score_grouped_device(sub, nq, d_a2.data_ptr(), d_p2.data_ptr(), 2, want.data_ptr(), 1.0, None, nodes=True)
        torch.cuda.synchronize(device)
        err, n_cmp = 0.0, 0
        for i in range(len(sub)):
            qb, qe = int(sub[i].query_begin), int(sub[i].query_end)
            err = max(err, float((got[qb:qe] - want[qb:qe]).abs().max())) if qe > qb else err
            n_cmp += qe - qb
        del full, fctx
        torch.cuda.empty_cache()
        err, = tm.max_over_ranks(err)
        assert err == 0.0, "%s scores differ from the unsharded single-GPU scores by %g" % (tables_mode, err)
        parity = {"queries_compared_per_rank": int(n_cmp), "parity_max_abs_err": err,
                  "against": "the same queries scored on one GPU holding the full (unsharded) tables"}
    if peers is not None:
        tm.barrier()
        peers.close()
    return {"wl": wl, "nq": nq, "dev_ms": dev_ms, "e2e_ms": e2e_ms, "loss": loss_ref, "launches": launches,
            "h2d": h2d, "call": call, "remote_rows": int(remote_rows), "formulas": len(wl.batches),
            "params": (tables, rels, pre, post), "fused_ms": fused_ms, "live_ms": live_ms,
            "live_launches": live_launches, "parity": parity, "lookup": lookup}


def measure_python_surface(args, tm, res, steps=20):
    """The headline mix through the Python operator surface: QueryEncoderDecoder.margin_loss(formula,
    queries) once per formula (what replaces model.py:112-127 as the reference's train/eval loops call
    it), negative sampling included.  (a) from StoreSlices of a QueryStore -- no per-query Python;
    (b) from lists of Query objects, the reference's own argument type (fewer steps: it is slow)."""
    import random
    import numpy as np
    import torch

    import graphqembed_b200 as gqe
    from graphqembed_b200.store import FormulaBlock
    from graphqembed_b200.synth import SynthKG

    wl, (tables, rels, pre, post) = res["wl"], res["params"]
    kg, d, device = wl.kg, wl.d, tm.device

    class G(object):
        pass
    graph = G()
    graph.relations, graph.features = kg.relations, res["lookup"]
    graph.full_lists = {m: kg.node_ids[m] for m in kg.modes}       # 1-chain negatives: any node of the mode
    dims = {m: d for m in kg.modes}
    feature_modules = {}
    for m, t in zip(kg.modes, tables):
        emb = torch.nn.Embedding(1, d)
        emb.weight = torch.nn.Parameter(t, requires_grad=False)     # the benchmark's own table, no copy
        feature_modules[m] = emb
    enc = gqe.get_encoder(0, graph, dims, feature_modules)
    dec = gqe.get_metapath_decoder(graph, dims, wl.decoder).to(device)
    idec = gqe.get_intersection_decoder(graph, dims, wl.inter).to(device)
    with torch.no_grad():
        for r, t in zip(kg.rel_keys, rels):
            dec.mats[r].copy_(t)
        for m, a, b in zip(kg.modes, pre, post):
            idec.pre_mats[m].copy_(a)
            idec.post_mats[m].copy_(b)
    model = gqe.QueryEncoderDecoder(graph, enc, dec, idec)
    model.negative_rng = np.random.default_rng(0)
    blocks = []
    for b in wl.batches:       # each formula's positives; negatives: 8 stored per query, drawn per step
        n = b.n_queries
        pairs = b.targets.reshape(-1, 2)
        negs = kg.sample_nodes(b.formula.target_mode, n * 8, np.random.RandomState(5)).astype(np.int32)
        ptr = np.arange(n + 1, dtype=np.int64) * 8
        blocks.append(FormulaBlock(b.formula, b.anchors.astype(np.int32), pairs[:, 0].astype(np.int32), ptr, negs,
                                   np.zeros(n + 1, dtype=np.int64), np.empty(0, dtype=np.int32)))

    def step_store():
        tot = 0.0
        with torch.no_grad():
            losses = [model.margin_loss(blk.formula, blk.all()) for blk in blocks]
        for blk, l in zip(blocks, losses):
            tot += float(l) * len(blk)
        return tot / wl.n_queries

    model.check_indices = False        # errors are polled once per step below, not once per formula
    for _ in range(3):
        step_store()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(steps):
        tm.flush.zero_()
        step_store()
        model.context().index_error()
    store_ms = (time.perf_counter() - t0) * 1e3 / steps

    def step_mix():
        with torch.no_grad():
            return float(model.margin_loss_mix([(blk.formula, blk.all()) for blk in blocks]))
    for _ in range(3):
        step_mix()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(steps):
        tm.flush.zero_()
        step_mix()
        model.context().index_error()
    mix_ms = (time.perf_counter() - t0) * 1e3 / steps
    # the same store uploaded once (QueryStore.to_device): batches sliced and negatives drawn on the GPU
    from graphqembed_b200.store import DeviceBlock
    dblocks = [DeviceBlock(blk, device) for blk in blocks]
    model.negative_seed = 2024

    def step_dev_store():
        with torch.no_grad():
            return float(model.margin_loss_mix([(blk.formula, blk.all()) for blk in dblocks]))
    for _ in range(3):
        step_dev_store()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(steps):
        tm.flush.zero_()
        torch.cuda.synchronize(device)
        step_dev_store()
        model.context().index_error()
    # (the flush + synchronise are inside this loop as in Timed.host_ms; subtract nothing: an upper bound)
    dev_store_ms = (time.perf_counter() - t0) * 1e3 / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        tm.flush.zero_()
        torch.cuda.synchronize(device)
    flush_ms = (time.perf_counter() - t0) * 1e3 / steps
    dev_store_ms -= flush_ms
    # the reference's training batch: ONE formula, 512 queries (train_helpers.py:95-107), host time per call
    small = dblocks[0].window(0, min(512, len(dblocks[0])))
    with torch.no_grad():
        for _ in range(20):
            model.margin_loss(small.formula, small)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        n_small = 300
        for _ in range(n_small):
            model.margin_loss(small.formula, small)
        small_host_us = (time.perf_counter() - t0) * 1e6 / n_small
        torch.cuda.synchronize(device)
        small_us = (time.perf_counter() - t0) * 1e6 / n_small
        small_h = blocks[0].window(0, len(small))
        for _ in range(20):
            model.margin_loss(small.formula, small_h)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(n_small):
            model.margin_loss(small.formula, small_h)
        torch.cuda.synchronize(device)
        small_h_us = (time.perf_counter() - t0) * 1e6 / n_small
    # lists of Query objects
    qlists = []
    for blk in blocks:
        s, rl = blk.formula.query_type, blk.formula.rels
        qlists.append([gqe.Query(SynthKG.query_graph(s, rl, int(blk.targets[i]), blk.anchors[:, i]),
                                 [int(x) for x in blk.negs[8 * i:8 * i + 8]], None, 100) for i in range(len(blk))])

    def step_objects():
        with torch.no_grad():
            return [float(model.margin_loss(blk.formula, qs)) for blk, qs in zip(blocks, qlists)]
    random.seed(0)
    step_objects()
    t0 = time.perf_counter()
    n_obj = max(2, steps // 8)
    for _ in range(n_obj):
        step_objects()
    obj_ms = (time.perf_counter() - t0) * 1e3 / n_obj
    return {"from_store": {"value": round(wl.n_queries / (store_ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(store_ms, 4),
                           "call": "QueryEncoderDecoder.margin_loss(formula, StoreSlice) x %d formulas: vectorised negative "
                                   "draw + pinned H2D of node ids + fused kernel + loss read" % len(blocks)},
            "from_store_one_call": {"value": round(wl.n_queries / (mix_ms * 1e-3), 1), "unit": UNIT,
                                    "ms_per_step": round(mix_ms, 4),
                                    "call": "QueryEncoderDecoder.margin_loss_mix([(formula, StoreSlice)] x %d): the same in ONE "
                                            "grouped launch" % len(blocks)},
            "from_device_store": {"value": round(wl.n_queries / (dev_store_ms * 1e-3), 1), "unit": UNIT,
                                  "ms_per_step": round(dev_store_ms, 4), "h2d_bytes_per_step": 0,
                                  "call": "QueryEncoderDecoder.margin_loss_mix([(formula, DeviceSlice)] x %d) on a store "
                                          "uploaded once (QueryStore.to_device): gqe_margin_loss_store_device = batch "
                                          "slicing + negative draw in one kernel, then the fused kernel; loss read "
                                          "back" % len(blocks),
                                  "batch_512": {"us_per_call_device_store": round(small_us, 1),
                                                "us_per_call_device_store_host_side": round(small_host_us, 1),
                                                "us_per_call_host_store": round(small_h_us, 1),
                                                "what": "margin_loss(formula, slice of 512 queries), back-to-back calls, "
                                                        "check_indices off: wall time per call incl. the final "
                                                        "synchronise / CPU time until the call returns / the same slice "
                                                        "as host arrays (vectorised draw + pinned H2D)"}},
            "from_query_objects": {"value": round(wl.n_queries / (obj_ms * 1e-3), 1), "unit": UNIT,
                                   "ms_per_step": round(obj_ms, 4),
                                   "call": "QueryEncoderDecoder.margin_loss(formula, [Query]) x %d formulas: the reference's "
                                           "argument type (random.choice per query, attribute walks)" % len(blocks)}}


def fused_kernel_span_ms(torch, ctx, step, tm, device, reps=8):
    """Duration of the fused tensor-core kernel ALONE, measured by the kernel itself: with the
    phase log on (gqe_debug_set_phase_log) thread 0 of every CTA stamps %globaltimer at entry and
    exit; the span first entry -> last exit of one launch, averaged over `reps` untimed steps (L2
    flushed like the timed ones).  None when the step does not run the tensor-core kernel in one
    launch.  CUDA events cannot isolate it: the kernel is launched from inside the C call, right
    behind gqe_pack (programmatic dependent launch)."""
    import numpy as np
    n_rec = 512                       # CTA records live in the last 256; tiles < 256 are stamped too
    log = torch.zeros(n_rec * 32, dtype=torch.int64, device=device)
    spans = []
    try:
        ctx.debug_set_phase_log(log.data_ptr(), n_rec)
        for _ in range(reps):
            log.zero_()
            tm.flush.zero_()
            step()
            torch.cuda.synchronize(device)
            rec = log.cpu().numpy().reshape(n_rec, 32)[::-1][:256]
            rec = rec[rec[:, 0] != 0]
            if len(rec) == 0:
                return None
            spans.append((rec[:, 8].max() - rec[:, 0].min()) * 1e-6)      # ns -> ms
    finally:
        ctx.debug_set_phase_log(None, 0)
    return float(np.mean(spans))


def measure_eval_shape(args, tm, local_rank, n_queries=8192, n_neg=100, d=256):
    """Scores only (no loss): every query of an intersection mix against 1 + n_neg targets."""
    import numpy as np
    import torch

    import graphqembed_b200 as gqe
    from graphqembed_b200 import _lib
    from graphqembed_b200.lowering import lower_formula
    from graphqembed_b200.query import Formula
    from graphqembed_b200.synth import bio_shaped
    from graphqembed_b200.workloads import Workload

    device, stream = tm.device, tm.stream
    kg = bio_shaped(seed=0)
    rng = np.random.RandomState(4242)
    T = n_neg + 1
    structures = ("2-inter", "3-inter", "3-inter_chain")
    wl = Workload("eval", kg, d, "bilinear", "mean", [])
    tables, rels, pre, post = device_parameters(wl, torch, device, seed=1234)
    lookup = gqe.RowLookup(kg.node_ids)
    mode_ids = {m: i for i, m in enumerate(kg.modes)}
    rel_ids = {r: i for i, r in enumerate(kg.rel_keys)}
    per = n_queries // len(structures)
    anchor_rows = np.zeros((_lib.GQE_MAX_ANCHORS, per * len(structures)), dtype=np.int32)
    target_rows = np.empty((per * len(structures), T), dtype=np.int32)
    items, n_rows = [], 0
    for i, s in enumerate(structures):
        f = Formula(s, kg.sample_rels(s, rng))
        b = kg.sample_batch(s, f.rels, per, n_neg, rng)
        q0 = i * per
        for k, mode in enumerate(f.anchor_modes):
            anchor_rows[k, q0:q0 + per] = lookup.rows(b["anchors"][k], mode)
        target_rows[q0:q0 + per, 0] = lookup.rows(b["target"], f.target_mode)
        target_rows[q0:q0 + per, 1:] = lookup.rows(b["negs"], f.target_mode)
        items.append((lower_formula(f, mode_ids, rel_ids), q0, q0 + per))
        n_rows += per * (len(f.anchor_modes) + T)
    nq = per * len(structures)
    segs = _lib.make_segments(items)
    ctx = gqe.Context(local_rank, stream.cuda_stream)
    ctx.bind_tables([t.data_ptr() for t in tables], [t.size(0) for t in tables], d)
    ctx.bind_relations(_lib.DECODER_ID["bilinear"], [r.data_ptr() for r in rels], d)
    ctx.bind_intersection(_lib.INTER_ID["mean"], [p.data_ptr() for p in pre], [p.data_ptr() for p in post], d)
    d_anchor = torch.from_numpy(anchor_rows).to(device)
    d_targets = torch.from_numpy(target_rows).to(device)
    d_scores = torch.empty(nq * T, dtype=torch.float32, device=device)

    def step():
        ctx.score_grouped_device(segs, nq, d_anchor.data_ptr(), d_targets.data_ptr(), T, d_scores.data_ptr(), 1.0, None)

    l0 = ctx.launch_count()
    step()
    launches = ctx.launch_count() - l0
    ms = tm.device_ms(step, max(10, args.steps // 4), 3)
    assert bool(torch.isfinite(d_scores).all())
    return {"ms": ms, "nq": nq, "T": T, "pairs": nq * T, "d": d, "launches": int(launches),
            "bytes": int(n_rows * (4 * d + 4) + nq * T * 4)}


def measure_train_step(args, tm, local_rank, batch=512, d=128, steps=30, cpu_steps=4):
    """BASELINE.json configs[0], as a TRAINING step (reference train_helpers.py:49-79): zero_grad,
    margin_loss on a 1-chain batch of 512, backward, Adam step over every parameter (dense table
    gradients, like the reference's nn.Embedding).  GPU: the drop-in modules (autograd.py + the VJP
    kernels); CPU: the oracle's torch ops with autograd on this box's host cores."""
    import random
    import numpy as np
    import torch

    import graphqembed_b200 as gqe
    from graphqembed_b200.synth import SynthKG, bio_shaped
    from oracle import netquery_oracle as O

    device = tm.device
    kg = bio_shaped(seed=0)
    rng = np.random.RandomState(7)
    rels = kg.sample_rels("1-chain", rng)
    b = kg.sample_batch("1-chain", rels, batch, 1, rng)

    class G(object):
        pass
    graph = G()
    graph.full_lists = kg.full_lists()
    graph.relations = kg.relations
    graph.features = gqe.RowLookup(kg.node_ids)
    dims = {m: d for m in kg.modes}
    torch.manual_seed(0)
    feature_modules = {m: torch.nn.Embedding(kg.sizes[m] + 2, d) for m in kg.modes}
    for m in kg.modes:
        feature_modules[m].weight.data.normal_(0, 1.0 / d)
    enc = gqe.get_encoder(0, graph, dims, feature_modules)
    dec = gqe.get_metapath_decoder(graph, dims, "bilinear")
    idec = gqe.get_intersection_decoder(graph, dims, "mean")
    model = gqe.QueryEncoderDecoder(graph, enc, dec, idec).to(device)
    f = gqe.Formula("1-chain", rels)
    qs = [gqe.Query(SynthKG.query_graph("1-chain", rels, b["target"][i], b["anchors"][:, i]), None, None)
          for i in range(batch)]

    def timed(model, opt):
        def step():
            opt.zero_grad()
            loss = model.margin_loss(f, qs)
            loss.backward()
            opt.step()
            return loss
        random.seed(0)
        for _ in range(5):
            step()
        torch.cuda.synchronize(device)
        l0 = model.context().launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = step()
        last = float(loss.item())
        torch.cuda.synchronize(device)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return ms, (model.context().launch_count() - l0) / steps, last

    import copy
    sparse_model = copy.deepcopy(model)
    native_model = copy.deepcopy(model)
    gpu_ms, launches, last = timed(model, torch.optim.Adam(model.parameters(), lr=0.01))
    sp_ms, sp_launches, sp_last = timed(sparse_model, gqe.SparseRowAdam(sparse_model, lr=0.01))

    # the same step as ONE native call (gqe_train_step_nodes_*): forward + backward + Adam launched
    # back to back from C++; indices as pinned int32 node ids (host call) or resident (device call)
    def native(nm, plan_of, batches, reps):
        """batches: [(formula, anchors int32 [A,n], pairs int32 [n,2])] -> (host ms, device ms, launches, loss)"""
        from graphqembed_b200 import _lib
        ctx = nm.context()
        hyper = _lib.AdamHyper(lr=0.01)
        items = []
        for f_, a_, p_ in batches:
            pin = torch.empty(a_.size + p_.size, dtype=torch.int32).pin_memory()
            pin.numpy()[:a_.size] = a_.reshape(-1)
            pin.numpy()[a_.size:] = p_.reshape(-1)
            dev = pin.to(device)
            items.append((plan_of(f_), p_.shape[0], pin, dev, a_.size))
        d_loss = torch.zeros(1, device=device)

        def host_pass():
            loss = 0.0
            for plan, n, pin, dev, off in items:
                loss = ctx.train_step_host(plan, n, pin.data_ptr(), pin.data_ptr() + 4 * off, 1.0, hyper, nodes=True)
            return loss

        def dev_pass():
            for plan, n, pin, dev, off in items:
                ctx.train_step_device(plan, n, dev.data_ptr(), dev.data_ptr() + 4 * off, 1.0, hyper, d_loss.data_ptr(),
                                      nodes=True)
        for _ in range(5):
            host_pass()
        torch.cuda.synchronize(device)
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        for _ in range(reps):
            loss = host_pass()
        host_ms = (time.perf_counter() - t0) * 1e3 / reps
        n_launch = (ctx.launch_count() - l0) / reps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev_pass()
        torch.cuda.synchronize(device)
        e0.record(torch.cuda.current_stream(device))
        for _ in range(reps):
            dev_pass()
        e1.record(torch.cuda.current_stream(device))
        torch.cuda.synchronize(device)
        return host_ms, e0.elapsed_time(e1) / reps, n_launch, float(loss)

    pairs0 = np.stack([b["target"], b["negs"][:, 0]], axis=1).astype(np.int32)
    nat_ms, nat_dev_ms, nat_launches, nat_last = native(native_model, native_model.plan,
                                                        [(f, b["anchors"].astype(np.int32), pairs0)], steps)
    # the headline mix as a training pass: one native step per formula of bio-mix-d256-b65536
    mix_res = None
    try:
        from graphqembed_b200 import _lib
        from graphqembed_b200.workloads import DEFAULT_WORKLOAD, WORKLOADS, make_workload
        wl = make_workload(DEFAULT_WORKLOAD, seed=0)
        mtables, mrels, mpre, mpost = device_parameters(wl, torch, device, seed=1234)
        mctx = gqe.Context(local_rank, torch.cuda.current_stream(device).cuda_stream)
        mctx.bind_tables([x.data_ptr() for x in mtables], [x.size(0) for x in mtables], wl.d)
        mctx.bind_relations(_lib.DECODER_ID[wl.decoder], [r.data_ptr() for r in mrels], wl.d)
        mctx.bind_intersection(_lib.INTER_ID[wl.inter], [x.data_ptr() for x in mpre], [x.data_ptr() for x in mpost], wl.d)
        lookup = gqe.RowLookup(wl.kg.node_ids)
        maps = lookup.device_maps(wl.kg.modes, [x.size(0) for x in mtables], device)
        mctx.bind_node_maps(maps[0], maps[1], maps[2])
        mode_ids = {m: i for i, m in enumerate(wl.kg.modes)}
        rel_ids = {r: i for i, r in enumerate(wl.kg.rel_keys)}

        class _M(object):
            def context(self):
                return mctx
        mb = [(bt.formula, np.ascontiguousarray(bt.anchors, dtype=np.int32),
               np.ascontiguousarray(bt.targets, dtype=np.int32).reshape(-1, 2)) for bt in wl.batches]
        h_ms, d_ms, n_l, l_last = native(_M(), lambda f_: gqe.lower_formula(f_, mode_ids, rel_ids), mb, 5)
        mix_res = {"workload": "%s as a training pass: one native step (forward + backward + Adam) per formula, %d formulas"
                               % (WORKLOADS[DEFAULT_WORKLOAD][0], len(mb)),
                   "ms_per_pass_host_call": round(h_ms, 3), "ms_per_pass_device": round(d_ms, 3),
                   "queries_per_s": round(wl.n_queries / d_ms * 1e3, 1), "gpu_kernels_per_pass": round(n_l, 1),
                   "loss_last_formula": l_last,
                   "note": "exact fp32 operator kernels (CUDA cores) with [d, n] intermediates in HBM: un-fused"}
        del mtables, mrels, mpre, mpost
    except Exception as exc:      # (never let the extra leg take the headline line down)
        mix_res = {"error": repr(exc)}

    # the reference's path on the host: same shapes, dense Adam over every tensor
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tables = {m: feature_modules[m].weight.detach().cpu().clone() for m in kg.modes}
    relp = {r: dec.mats[r].detach().cpu().clone() for r in kg.rel_keys}
    pre = {m: idec.pre_mats[m].detach().cpu().clone() for m in kg.modes}
    post = {m: idec.post_mats[m].detach().cpu().clone() for m in kg.modes}
    orc = O.OracleScorer(tables, kg.node_maps(), relp, "bilinear", "mean", pre, post, full_lists=kg.full_lists())
    leaves = list(orc.tables.values()) + list(orc.rel_params.values()) + list(orc.pre.values()) + list(orc.post.values())
    for t in leaves:
        t.requires_grad_(True)
    copt = torch.optim.Adam(leaves, lr=0.01)
    of = O.Formula("1-chain", rels)
    oqs = [O.Query(SynthKG.query_graph("1-chain", rels, b["target"][i], b["anchors"][:, i]), None, None)
           for i in range(batch)]

    def cpu_step():
        copt.zero_grad()
        loss = orc.margin_loss(of, oqs)
        loss.backward()
        copt.step()

    cpu_step()
    t0 = time.perf_counter()
    for _ in range(cpu_steps):
        cpu_step()
    cpu_ms = (time.perf_counter() - t0) * 1e3 / cpu_steps
    return {"workload": "Bio KG 1-chain, Bilinear, d=%d, batch=%d: zero_grad + margin_loss + backward + Adam step "
                        "(dense table gradients)" % (d, batch),
            "gpu_ms_per_step": round(gpu_ms, 3), "gpu_queries_per_s": round(batch / gpu_ms * 1e3, 1),
            "gpu_kernels_per_step": round(launches, 1), "loss_after": last,
            "sparse": {"what": "same step with SparseRowAdam: table gradients as (row, gradient) pairs, row-wise Adam "
                               "with exact catch-up of the zero-gradient steps (dense-Adam trajectory)",
                       "gpu_ms_per_step": round(sp_ms, 3), "gpu_queries_per_s": round(batch / sp_ms * 1e3, 1),
                       "gpu_kernels_per_step": round(sp_launches, 1), "loss_after": sp_last},
            "native": {"what": "same step as ONE native call (gqe_train_step_nodes_host / _device): forward + backward + "
                               "row-wise Adam with exact catch-up, no Python between the kernels",
                       "gpu_ms_per_step": round(nat_ms, 4), "device_ms_per_step": round(nat_dev_ms, 4),
                       "gpu_queries_per_s": round(batch / nat_ms * 1e3, 1), "gpu_kernels_per_step": round(nat_launches, 1),
                       "loss_after": nat_last},
            "native_mix": mix_res,
            "cpu_ms_per_step": round(cpu_ms, 3), "cpu_queries_per_s": round(batch / cpu_ms * 1e3, 1), "cpu_cores": cores,
            "note": "wall clock per step incl. Python; CPU = oracle port with torch autograd + torch.optim.Adam"}


def ncu_record(name):
    """Numbers copied from the committed ncu capture of this workload's dominant kernel
    (profiles/traffic.json: DRAM bytes per launch, tensor-pipe activity, kernel duration under ncu)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return {}
    with open(tpath) as fh:
        rec = json.load(fh).get(name)
    if isinstance(rec, (int, float)):
        rec = {"bytes": rec}
    return rec or {}


def roofline_of(wl, name, ms_per_step, pk, fused_ms=None, live_ms=None):
    """Roofline of the dominant kernel for one step of `wl` on one GPU.

    Bilinear / DeepSets workloads are bound by the tensor pipe: the d x d contractions run as three
    bf16 tensor-core products per algorithmic one (hi*hi + lo*hi + hi*lo, fp32 accumulate), so
    `achieved` = 3 x algorithmic flops / kernel time, against the measured BURST bf16 peak (the kernel
    is one ~0.1 ms launch timed alone behind an L2 flush).  `frac_issued` counts only the flops the
    tensor pipe really executed after runs of linear operators were pre-multiplied.  With the
    weights cached the timed step IS the fused kernel (one launch between the CUDA events).
    TransE / DistMult + element-wise intersections have no contraction: HBM-gather-bound."""
    bytes_alg, flops_alg = wl.algorithmic_bytes(), wl.algorithmic_flops()
    sec = ms_per_step * 1e-3
    gbs = bytes_alg / sec / 1e9
    passes = 3
    tfl_exec = passes * flops_alg / sec / 1e12
    t_hbm = bytes_alg / (pk["hbm_gbs"] * 1e9)
    t_tc = passes * flops_alg / (pk["bf16_tflops"] * 1e12)
    rec = ncu_record(name)
    bound = "hbm" if t_hbm >= t_tc else "tensor"
    out = {"bound": bound,
           "achieved": round(gbs if bound == "hbm" else tfl_exec, 3),
           "peak": pk["hbm_gbs"] if bound == "hbm" else pk["bf16_tflops"],
           "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
           "frac": round((gbs / pk["hbm_gbs"]) if bound == "hbm" else (tfl_exec / pk["bf16_tflops"]), 4),
           "traffic": rec.get("bytes"), "peak_source": pk["source"],
           "peak_kind": "measured copy bandwidth" if bound == "hbm" else "measured bf16 dense, burst (kernel timed alone)",
           "kernel_ms": round(ms_per_step, 5),
           "algorithmic_bytes_per_launch": bytes_alg, "algorithmic_flops_per_launch": flops_alg,
           "hbm": {"achieved_gbs": round(gbs, 2), "frac": round(gbs / pk["hbm_gbs"], 4)}}
    if wl.decoder == "bilinear":
        issued = passes * wl.composed_flops(min_rows=0) / sec / 1e12
        out["kernel"] = ("gqe_fused_tc<%d,%s> (persistent tcgen05 kernel; with the packed weights cached it is the "
                         "only launch of the step)" % (wl.d, "-1" if len(wl.batches) > 1 else "structure"))
        out["frac_algorithmic"] = round(tfl_exec / pk["bf16_tflops"], 4)
        out["frac_issued"] = round(issued / pk["bf16_tflops"], 4)
        out["tensor"] = {"executed_tflops": round(tfl_exec, 3), "algorithmic_tflops": round(flops_alg / sec / 1e12, 3),
                         "issued_tflops": round(issued, 3), "peak_tflops": pk["bf16_tflops"],
                         "peak_tflops_sustained": pk["bf16_tflops_sustained"],
                         "tensor_pipe_active_pct_ncu": rec.get("tensor_pipe_active_pct"),
                         "kernel_us_under_ncu": rec.get("kernel_us"),
                         "note": "executed = 3 x algorithmic (split-bf16: hi*hi + lo*hi + hi*lo); issued = what the "
                                 "tensor pipe ran after pre-composition; tensor_pipe_active_pct_ncu = "
                                 "sm__pipe_tensor_cycles_active of the committed ncu capture (profiles/)"}
        if fused_ms:
            out["fused_kernel_span"] = {"ms": round(fused_ms, 5),
                                        "how": "first CTA entry -> last CTA exit, %globaltimer stamped by the kernel itself"}
        if live_ms:
            out["weights_live_ms"] = round(live_ms, 5)
    else:
        out["kernel"] = "gqe_fused_vec<%d> (streaming warp-per-query kernel, no contraction)" % wl.d
    return out


NVLINK_PEER_GBS = 770.0     # /opt/skills/guides/B200_PROFILING.md: measured peer copy, per direction per GPU


def summarize(name, r, pk, world, WORKLOADS):
    """One secondary workload as a JSON object."""
    wl = r["wl"]
    sec = r["dev_ms"] * 1e-3
    out = {"workload": WORKLOADS[name][0], "name": name, "queries_per_step_per_gpu": r["nq"], "formulas": r["formulas"],
           "decoder": wl.decoder, "intersection": wl.inter, "d": wl.d,
           "value": round(world * r["nq"] / sec, 1), "unit": UNIT, "ms_per_step": round(r["dev_ms"], 5),
           "e2e": {"value": round(world * r["nq"] / (r["e2e_ms"] * 1e-3), 1), "ms_per_step": round(r["e2e_ms"], 5),
                   "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4},
           "gpu_launches_per_step": round(r["launches"], 2), "loss": r["loss"],
           "roofline": roofline_of(wl, name, r["dev_ms"], pk, r.get("fused_ms"), r.get("live_ms"))}
    if r.get("live_ms"):
        out["weights_live"] = {"ms_per_step": round(r["live_ms"], 5), "value": round(world * r["nq"] / (r["live_ms"] * 1e-3), 1),
                               "gpu_launches_per_step": round(r["live_launches"], 2)}
    if r.get("parity"):
        out["parity"] = r["parity"]
    return out


def run_native(args):
    import torch

    from graphqembed_b200.workloads import DEFAULT_WORKLOAD, LARGE_WORKLOAD, WORKLOADS

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    stream = torch.cuda.current_stream(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2
    tm = Timed(torch, dist, device, stream, flush)

    name = args.workload or DEFAULT_WORKLOAD
    tables_mode = args.tables or ("p2p" if name.startswith("synth-10m") and world > 1 else "replicated")
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = measure(args, name, tables_mode, tm, rank, world, local_rank, sampler, args.formulas_per_structure)
    sampler.stop()
    surface = None
    if world == 1 and rank == 0 and not args.no_extras and WORKLOADS[name][4] == "bilinear":
        surface = measure_python_surface(args, tm, res)

    # BASELINE.json configs[4]: the node-type-sharded 10 M-node table, measured beside the headline
    # whenever the job has more than one GPU; every rank checks its scores against the unsharded tables
    sharded_res = {}
    if world > 1 and args.workload is None and not args.no_sharded:
        for key, mode, routed in (("p2p", "p2p", True), ("p2p_unrouted", "p2p", False), ("staged", "staged", True)):
            r = measure(args, LARGE_WORKLOAD, mode, tm, rank, world, local_rank, None, args.sharded_formulas,
                        steps=max(10, args.steps // 2), parity_slice=4096 // (6 * args.sharded_formulas) + 1,
                        want_live=False, routed=routed)
            r.pop("params")
            r["routed"] = routed
            sharded_res[key] = r
            torch.cuda.empty_cache()

    # every other BASELINE config + the HBM-bound operators, on one GPU
    extras, hbm = {}, {}
    if world == 1 and args.workload is None and not args.no_extras:
        plan = [("configs[0]", "bio-edge-d128-b512", 1), ("configs[1]", "bio-chain-d128-b4096", 1),
                ("configs[2]", "bio-inter-d128-b8192", 1), ("configs[2]-min", "bio-inter-d128-b8192-min", 1),
                ("configs[3]-66-formulas", "bio-mix-d256-b65536", 11),
                ("configs[4]-1gpu", LARGE_WORKLOAD, args.sharded_formulas)]
        for key, wname, fps in plan:
            r = measure(args, wname, "replicated", tm, rank, world, local_rank, None, fps, steps=args.extra_steps)
            r.pop("params")
            extras[key] = r
            extras[key]["name"] = wname
            torch.cuda.empty_cache()
        for wname in ("synth-10m-transe-minsimple-d256-b65536", "synth-10m-distmult-meansimple-d256-b65536",
                      "bio-transe-minsimple-d256-b65536"):
            r = measure(args, wname, "replicated", tm, rank, world, local_rank, None, args.sharded_formulas,
                        steps=args.extra_steps, want_live=False)
            r.pop("params")
            hbm[wname] = r
            torch.cuda.empty_cache()

    # the evaluation shape (utils.py:70-91): one query against its positive + 100 negatives
    eval_res = None
    if world == 1 and not args.no_eval_shape:
        eval_res = measure_eval_shape(args, tm, local_rank)

    train_res = None
    if world == 1 and not args.no_train_step:
        train_res = measure_train_step(args, tm, local_rank)

    line = None
    if rank == 0:
        pk = peaks()
        wl, nq = res["wl"], res["nq"]
        ms_per_step = res["dev_ms"]
        line = {
            "metric": METRIC, "value": round(world * nq / (ms_per_step * 1e-3), 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[name][0], "name": name, "queries_per_step_per_gpu": nq,
                       "targets_per_query": 2, "decoder": wl.decoder, "intersection": wl.inter, "d": wl.d,
                       "formulas": res["formulas"],
                       "indices": "int32 NODE IDS; node id -> table row (node_maps[mode][n] + 1) inside the kernel",
                       "contractions": "bf16x3 split on tcgen05, fp32 accumulate (scores within 1e-4 of fp32)",
                       "weights": "packed / pre-multiplied operator matrices cached across steps (parameters unchanged "
                                  "between steps; `weights_live` = re-prepared every step)",
                       "tables": {"replicated": "replicated per GPU (Bio-size); queries sharded, no collective",
                                  "p2p": "sharded by node type; remote rows read over NVLink (CUDA IPC peer mappings) inside the fused kernel: "
                                         "fetched a tile ahead by its TMA helper warp into a local staging area "
                                         "(GQE_STAGE=0: gathered in place)",
                                  "staged": "sharded by node type; NCCL all-to-all row exchange"}[tables_mode],
                       "l2": "flushed before every step (256 MiB write)"},
            "e2e": {"value": round(world * nq / (res["e2e_ms"] * 1e-3), 1), "unit": UNIT,
                    "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": 4,
                    "ms_per_step": round(res["e2e_ms"], 5), "call": res["call"]},
            "gpu_launches": int(round(res["launches"] * args.steps)),
            "gpu_launches_per_step": round(res["launches"], 2),
            "clocks": sampler.summary(),
            "roofline": roofline_of(wl, name, ms_per_step, pk, res.get("fused_ms"), res.get("live_ms")),
            "loss": res["loss"],
        }
        if res.get("live_ms"):
            line["weights_live"] = {"ms_per_step": round(res["live_ms"], 5),
                                    "value": round(world * nq / (res["live_ms"] * 1e-3), 1),
                                    "gpu_launches_per_step": round(res["live_launches"], 2),
                                    "note": "weight cache off: gqe_compose + gqe_pack run in front of the fused kernel "
                                            "every step, as when an optimiser changes the matrices every step"}
        if surface:
            line["e2e"].update(surface)
        for mode, r in sharded_res.items():
            sec = r["dev_ms"] * 1e-3
            nv_bytes = r["remote_rows"] * r["wl"].d * 4
            line.setdefault("sharded", {})[mode] = {
                "workload": WORKLOADS[LARGE_WORKLOAD][0], "value": round(world * r["nq"] / sec, 1), "unit": UNIT,
                "per_gpu": round(r["nq"] / sec, 1),
                "ms_per_step": round(r["dev_ms"], 5), "e2e_value": round(world * r["nq"] / (r["e2e_ms"] * 1e-3), 1),
                "gpu_launches_per_step": round(r["launches"], 2), "formulas": r["formulas"], "loss_rank0": r["loss"],
                "parity": r["parity"],
                "queries": ("routed: every rank scores the formulas whose target node type it owns "
                            "(sharded.route_by_target_mode)" if r["routed"] else
                            "unrouted: formulas with arbitrary target node types on every rank"),
                "remote_row_fraction_rank0": round(r["remote_rows"] / float(sum(
                    b.n_queries * (len(b.formula.anchor_modes) + 2) for b in r["wl"].batches)), 4),
                "hbm": {"achieved_gbs": round(r["wl"].algorithmic_bytes() / sec / 1e9, 2),
                        "frac": round(r["wl"].algorithmic_bytes() / sec / 1e9 / pk["hbm_gbs"], 4)},
                "nvlink": {"remote_rows_per_step_rank0": r["remote_rows"], "bytes_per_step_rank0": nv_bytes,
                           "achieved_gbs_in": round(nv_bytes / sec / 1e9, 2), "peak_gbs": NVLINK_PEER_GBS,
                           "frac": round(nv_bytes / sec / 1e9 / NVLINK_PEER_GBS, 4),
                           "note": "inbound row bytes of rank 0 / step time vs the measured peer-copy bandwidth"}}
        if extras:
            line["configs"] = {k: summarize(r["name"], r, pk, world, WORKLOADS) for k, r in extras.items()}
        if hbm:
            line["hbm_bound"] = {k: summarize(k, r, pk, world, WORKLOADS) for k, r in hbm.items()}
        if eval_res is not None:
            sec = eval_res["ms"] * 1e-3
            line["eval_shape"] = {
                "workload": "Bio KG 2/3-inter + 3-inter_chain, d=%d, %d queries x (1 positive + %d negatives)" % (
                    eval_res["d"], eval_res["nq"], eval_res["T"] - 1),
                "pairs_per_s": round(eval_res["pairs"] / sec, 1), "queries_per_s": round(eval_res["nq"] / sec, 1),
                "ms_per_step": round(eval_res["ms"], 5), "gpu_launches_per_step": eval_res["launches"],
                "hbm": {"algorithmic_bytes": eval_res["bytes"], "achieved_gbs": round(eval_res["bytes"] / sec / 1e9, 2),
                        "frac": round(eval_res["bytes"] / sec / 1e9 / pk["hbm_gbs"], 4),
                        "note": "gather-bound: one table row per (query, target) pair; Bio-size tables are "
                                "L2-resident after first touch, so this can exceed the DRAM roofline"}}
        if train_res is not None:
            line["train_step"] = train_res
        if world == 1 and not args.no_cpu_baseline:
            tables, rels, pre, post = res["params"]
            line["cpu_baseline"] = cpu_reference(args, name, tables=[t.cpu() for t in tables],
                                                 rels=[r.cpu() for r in rels], pre=[p.cpu() for p in pre],
                                                 post=[p.cpu() for p in post], steps=None)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def cpu_reference(args, name, tables=None, rels=None, pre=None, post=None, steps=None, warmup=2):
    """Time the oracle port of the reference path on this box's host cores on a
    bounded sample of the workload.  Includes the reference's own Python index
    building (list comprehensions + node_maps dict lookups): that is its real
    path (model.py:75-92, bio/data_utils.py:20-21)."""
    import numpy as np
    import torch

    from graphqembed_b200.synth import SynthKG
    from graphqembed_b200.workloads import WORKLOADS, make_workload
    from oracle import netquery_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(args.cpu_sample, WORKLOADS[name][2])
    wl = make_workload(name, seed=0, total=sample, formulas_per_structure=args.formulas_per_structure)
    kg = wl.kg
    if tables is None:
        import math
        g = torch.Generator().manual_seed(1234)
        d = wl.d
        bound = math.sqrt(6.0 / (2 * d))
        uni = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * bound
        tables = [torch.randn(kg.sizes[m] + 2, d, generator=g) * (1.0 / d) for m in kg.modes]
        rels = [uni(d, d) for _ in kg.rel_keys]
        pre = [uni(d, d) for _ in kg.modes]
        post = [uni(d, d) for _ in kg.modes]
    orc = O.OracleScorer(dict(zip(kg.modes, tables)), kg.node_maps(), dict(zip(kg.rel_keys, rels)), wl.decoder,
                         wl.inter, dict(zip(kg.modes, pre)), dict(zip(kg.modes, post)))
    work = []
    for b in wl.batches:
        s = b.formula.query_type
        f = O.Formula(s, b.formula.rels)
        pairs = b.targets.reshape(-1, 2)
        qs = [O.Query(SynthKG.query_graph(s, f.rels, pairs[i, 0], b.anchors[:, i]), None, None)
              for i in range(b.n_queries)]
        work.append((f, qs, [int(x) for x in pairs[:, 1]]))

    def one_pass():
        tot = 0.0
        with torch.no_grad():
            for f, qs, negs in work:
                tot += float(orc.margin_loss(f, qs, neg_nodes=negs)) * len(qs)
        return tot / wl.n_queries

    for _ in range(warmup):
        one_pass()
    times = []
    if steps is None:
        t_end = time.perf_counter() + args.cpu_seconds
        while time.perf_counter() < t_end or len(times) < 3:
            t0 = time.perf_counter()
            one_pass()
            times.append(time.perf_counter() - t0)
    else:
        for _ in range(steps):
            t0 = time.perf_counter()
            one_pass()
            times.append(time.perf_counter() - t0)
    mean_s = sum(times) / len(times)
    return {"value": round(wl.n_queries / mean_s, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d queries of %s (same mix), %d passes, torch %s CPU, %d threads; oracle/netquery_oracle.py "
                      "margin_loss incl. Python index building" % (wl.n_queries, name, len(times), torch.__version__,
                                                                    cores),
            "ms_per_pass": round(mean_s * 1e3, 3)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.
    The reference is Python and cannot travel to the GPU box, so this arm times
    its oracle port (bit-identical to the reference, see tests/) with all host
    threads; each step is one pass over a bounded sample of the workload."""
    from graphqembed_b200.workloads import DEFAULT_WORKLOAD, WORKLOADS
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    base = cpu_reference(args, name, steps=args.steps, warmup=max(args.warmup, 1))
    desc, d, total, structures, decoder, inter = WORKLOADS[name]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": base["ms_per_pass"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": name, "targets_per_query": 2, "decoder": decoder, "intersection": inter,
                   "d": d, "sample_queries_per_step": min(args.cpu_sample, total)},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
