"""Node-type-sharded tables on real GPUs: both data paths must give the SAME
BITS as the replicated single-GPU call (the same fp32 rows reach the same
kernel arithmetic), and that call is itself checked against the oracle in
test_gpu_parity.py.  Uses 2 GPUs when the box has them (CUDA IPC peer mapping
+ NCCL exchange), else 1 (the staged path's gather / staging / index rewrite
with a world of one)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, d, ret):
    import torch.distributed as dist

    import graphqembed_b200 as gqe
    from graphqembed_b200 import _lib, sharded
    from graphqembed_b200.synth import SynthKG
    from graphqembed_b200.workloads import Workload
    from graphqembed_b200.query import Formula, QueryBatch

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        kg = SynthKG(["a", "b", "c"], [700, 500, 900], n_rel_pairs=6, seed=5, self_loops=False)
        rng = np.random.RandomState(50 + rank)
        batches = []
        # d = 256: enough tiles (> 2 per SM) that the CTAs of the grouped kernel work through several each -- the
        # helper warp of its STAGE instantiation then fetches the peer rows of every tile but a CTA's first
        scale = 48 if d == 256 else 1
        for s, n in (("1-chain", 130), ("3-chain", 77), ("2-inter", 200), ("3-inter", 129), ("3-inter_chain", 64),
                     ("3-chain_inter", 31)):
            rels = kg.sample_rels(s, rng)
            b = kg.sample_batch(s, rels, n * scale, 1, rng)
            batches.append(QueryBatch(Formula(s, rels), b["anchors"], np.stack([b["target"], b["negs"][:, 0]], 1).reshape(-1)))
        wl = Workload("toy", kg, d, "bilinear", "mean", batches)
        g = torch.Generator(device=device).manual_seed(99)
        full = [torch.randn(kg.sizes[m] + 2, d, generator=g, device=device) / d for m in kg.modes]
        mats = lambda n: [(torch.rand(d, d, generator=g, device=device) - 0.5) * 0.3 for _ in range(n)]
        rels, pre, post = mats(len(kg.rel_keys)), mats(3), mats(3)
        lookup = gqe.RowLookup(kg.node_ids)
        mode_ids = {m: i for i, m in enumerate(kg.modes)}
        rel_ids = {r: i for i, r in enumerate(kg.rel_keys)}
        segs, anchor_rows, pair_rows = wl.lower(lookup, mode_ids, rel_ids)
        nq = wl.n_queries
        rows = [t.size(0) for t in full]

        def new_ctx(ptrs, nrows):
            c = gqe.Context(rank, torch.cuda.current_stream().cuda_stream)
            c.bind_relations(0, [r.data_ptr() for r in rels], d)
            c.bind_intersection(0, [p.data_ptr() for p in pre], [p.data_ptr() for p in post], d)
            c.bind_tables(ptrs, nrows, d)
            return c

        def run(c, a, p):
            scores = torch.empty(nq, 2, device=device)
            loss = torch.empty(1, device=device)
            c.score_grouped_device(segs, nq, a.data_ptr(), p.data_ptr(), 2, scores.data_ptr(), 1.0, loss.data_ptr())
            torch.cuda.synchronize()
            return scores.cpu(), float(loss.item())

        d_anchor, d_pairs = torch.from_numpy(anchor_rows).to(device), torch.from_numpy(pair_rows).to(device)
        want_scores, want_loss = run(new_ctx([t.data_ptr() for t in full], rows), d_anchor, d_pairs)

        owner = sharded.owner_by_node_type(3, world)
        own_ptrs = [full[m].data_ptr() if owner[m] == rank else 0 for m in range(3)]
        own_rows = [rows[m] if owner[m] == rank else 0 for m in range(3)]
        ok = {}

        # a plan that needs a shard this rank does not hold must fail loudly, not read garbage
        if world > 1:
            try:
                run(new_ctx(own_ptrs, own_rows), d_anchor, d_pairs)
                ok["unbound"] = False
            except gqe.GqeError as e:
                ok["unbound"] = e.code == -2

        # staged: NCCL request/row all-to-all + owner gather kernel + staging tables
        octx = new_ctx(own_ptrs, own_rows)
        info = sharded.chunks_of_segments(segs, nq, 2)
        plan = sharded.ExchangePlan(owner, world, rank, [(m, n) for m, n, _, _, _ in info])
        ex = sharded.RowExchange(plan, d, sharded.device_gather(octx), device=device)
        req, s_anchor, s_pairs = sharded.stage_grouped(plan, info, segs, anchor_rows, pair_rows, 2)
        ex.run(torch.from_numpy(req).to(device))
        st_ptrs, st_rows = ex.staging_tables()
        got_scores, got_loss = run(new_ctx(st_ptrs, st_rows), torch.from_numpy(s_anchor).to(device),
                                   torch.from_numpy(s_pairs).to(device))
        ok["staged"] = torch.equal(got_scores, want_scores) and got_loss == want_loss

        # in place over NVLink: peers mapped with CUDA IPC
        if world > 1:
            peers = sharded.PeerTables(octx, owner, rows, {m: full[m] for m in range(3) if owner[m] == rank})
            ptrs = peers.pointers()
            ok["peer_ptrs_differ"] = all((ptrs[m] == full[m].data_ptr()) == (owner[m] == rank) for m in range(3))
            got_scores, got_loss = run(new_ctx(ptrs, rows), d_anchor, d_pairs)
            ok["p2p"] = torch.equal(got_scores, want_scores) and got_loss == want_loss
            dist.barrier()
            peers.close()
        ret[rank] = ok
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("d", [128, 256])
def test_sharded_paths_bit_identical_to_replicated(d):
    import torch.multiprocessing as mp
    world = 2 if torch.cuda.device_count() >= 2 else 1
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, d, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = dict(ret)
    assert sorted(got) == list(range(world))
    for r in range(world):
        assert all(got[r].values()), (r, got[r])
        assert "staged" in got[r] and (world == 1 or "p2p" in got[r])
