"""Host logic of the node-type-sharded tables (graphqembed_b200/sharded.py):
ownership, request ordering, split sizes and index rewriting, exercised over a
REAL world_size-2 process group (gloo, CPU tensors).  The row gather itself is
a CUDA kernel in the product; here the test supplies a torch-CPU stand-in as
the checker, so only the exchange plumbing is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphqembed_b200 import _lib
from graphqembed_b200.sharded import (ExchangePlan, RowExchange, chunks_of_segments, owner_by_node_type,
                                      stage_grouped)

N_MODES, D, ROWS = 3, 8, [50, 70, 30]
# (structure, target_mode, anchor modes) of the toy segments
TOY = [(0, 0, (1,)), (3, 1, (0, 2)), (4, 2, (0, 1, 2)), (5, 0, (2, 1))]


def full_tables():
    g = torch.Generator().manual_seed(7)
    return [torch.randn(r, D, generator=g) for r in ROWS]


def toy_batch(rank, T=2):
    rng = np.random.RandomState(100 + rank)
    items, q0 = [], 0
    sizes = [5 + rank, 9, 0 if rank == 0 else 4, 7]       # includes an EMPTY segment on rank 0
    for (structure, tm, ams), n in zip(TOY, sizes):
        pl = _lib.Plan()
        pl.structure, pl.target_mode, pl.inter_mode = structure, tm, -1
        for k in range(3):
            pl.anchor_mode[k] = ams[k] if k < len(ams) else -1
            pl.rel[k] = 0
        items.append((pl, q0, q0 + n))
        q0 += n
    segs = _lib.make_segments(items)
    anchor = np.zeros((3, q0), dtype=np.int32)
    target = np.zeros((q0, T), dtype=np.int32)
    for (structure, tm, ams), (pl, b, e) in zip(TOY, items):
        target[b:e] = rng.randint(0, ROWS[tm], size=(e - b, T))
        for k, m in enumerate(ams):
            anchor[k, b:e] = rng.randint(0, ROWS[m], size=e - b)
    return segs, q0, anchor, target


def test_plan_layout_single_rank():
    owner = owner_by_node_type(N_MODES, 1)
    segs, nq, anchor, target = toy_batch(1)
    info = chunks_of_segments(segs, nq, 2)
    plan = ExchangePlan(owner, 1, 0, [(m, n) for m, n, _, _, _ in info])
    plan.set_counts([plan.mode_count])
    req, s_anchor, s_target = stage_grouped(plan, info, segs, anchor, target, 2)
    assert plan.n_send == plan.n_recv == req.size == sum(n for _, n, _, _, _ in info)
    # the request vector, cut per mode, IS the staging table's row list
    tabs = full_tables()
    for c, (m, n, kind, si, k) in enumerate(info):
        b, e = int(segs[si].query_begin), int(segs[si].query_end)
        orig = target.reshape(-1)[2 * b:2 * e] if kind == "target" else anchor[k, b:e]
        stag = s_target.reshape(-1)[2 * b:2 * e] if kind == "target" else s_anchor[k, b:e]
        block = req[plan.mode_offset[m]:plan.mode_offset[m] + plan.mode_count[m]]
        assert np.array_equal(block[stag], orig)
        assert np.array_equal(tabs[m][block.astype(np.int64)][stag.astype(np.int64)].numpy(), tabs[m][orig.astype(np.int64)].numpy())


def test_plan_rejects_inconsistent_counts():
    plan = ExchangePlan([0, 1], 2, 0, [(0, 3), (1, 2)])
    with pytest.raises(ValueError):
        plan.set_counts([[3, 9], [0, 0]])
    with pytest.raises(ValueError):
        ExchangePlan([0, 1], 2, 0, [(5, 1)])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tabs = full_tables()
        owner = owner_by_node_type(N_MODES, world)         # modes 0,2 -> rank 0; mode 1 -> rank 1
        served = []

        def gather(mode_id, rows, out):                     # test stand-in for the CUDA gather kernel
            assert owner[mode_id] == rank, "asked to serve a shard this rank does not own"
            served.append((mode_id, rows.numel()))
            out.copy_(tabs[mode_id][rows.long()])

        segs, nq, anchor, target = toy_batch(rank)
        info = chunks_of_segments(segs, nq, 2)
        plan = ExchangePlan(owner, world, rank, [(m, n) for m, n, _, _, _ in info])
        ex = RowExchange(plan, D, gather)
        req, s_anchor, s_target = stage_grouped(plan, info, segs, anchor, target, 2)
        for _ in range(2):                                  # buffers are reusable across steps
            rows = ex.run(torch.from_numpy(req))
        ptrs, counts = ex.staging_tables()
        ok = True
        for c, (m, n, kind, si, k) in enumerate(info):
            b, e = int(segs[si].query_begin), int(segs[si].query_end)
            orig = target.reshape(-1)[2 * b:2 * e] if kind == "target" else anchor[k, b:e]
            stag = s_target.reshape(-1)[2 * b:2 * e] if kind == "target" else s_anchor[k, b:e]
            staging = rows[int(plan.mode_offset[m]):int(plan.mode_offset[m] + plan.mode_count[m])]
            ok = ok and torch.equal(staging[torch.from_numpy(stag).long()], tabs[m][torch.from_numpy(orig).long()])
            ok = ok and (ptrs[m] == rows.data_ptr() + int(plan.mode_offset[m]) * D * 4)
        ok = ok and counts == [int(x) for x in plan.mode_count]
        ok = ok and all(owner[m] == rank for m, _ in served)
        ok = ok and sum(n for _, n in served) == 2 * plan.n_recv
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_row_exchange_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_route_by_target_mode_keeps_targets_local():
    """sharded.route_by_target_mode: a formula is scored where its target node type lives."""
    from graphqembed_b200 import sharded
    from graphqembed_b200.query import Formula
    from graphqembed_b200.synth import STRUCTURES, bio_shaped
    kg = bio_shaped(scale=0.01)
    rng = np.random.RandomState(0)
    formulas = [Formula(s, kg.sample_rels(s, rng)) for s in STRUCTURES for _ in range(5)]
    mode_ids = {m: i for i, m in enumerate(kg.modes)}
    for world in (1, 2, 4, 8):
        owner = sharded.owner_by_node_type(len(kg.modes), world)
        routed = sharded.route_by_target_mode(formulas, mode_ids, owner)
        assert sum(len(v) for v in routed.values()) == len(formulas)
        for rank, fs in routed.items():
            assert all(owner[mode_ids[f.target_mode]] == rank for f in fs)
            assert set(mode_ids[f.target_mode] for f in fs) <= set(sharded.modes_owned_by(owner, rank))
    # the workload generator can draw a rank's formulas from the node types it owns
    from graphqembed_b200.workloads import make_workload
    wl = make_workload("bio-mix-d256-b65536", kg=kg, total=60, target_modes=[kg.modes[1], kg.modes[3]])
    assert set(b.formula.target_mode for b in wl.batches) <= {kg.modes[1], kg.modes[3]}
