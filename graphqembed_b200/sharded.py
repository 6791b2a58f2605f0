"""Node-type-sharded embedding tables across the GPUs of one box.

The reference is single-process: every lookup goes to one ``nn.Embedding`` per
mode in one address space (reference netquery/bio/data_utils.py:16-21).  Here
the tables are partitioned BY NODE TYPE over the ranks of a
``torch.distributed`` group (one process per GPU); the query batch is split
data-parallel, and only the rows a rank needs from shards it does not own cross
NVLink.  Two data paths, bit-identical in their results (the same fp32 rows
reach the same kernel arithmetic):

``PeerTables`` (the product path)
    every rank maps every other rank's shard with CUDA IPC
    (``gqe_ipc_export`` / ``gqe_ipc_open``) and binds the peer pointers as
    ordinary tables; the fused scoring kernel then gathers remote rows IN
    PLACE over NVLink/NVSwitch -- no staging buffer, no collective, the
    transfer overlaps the tensor-core work of the other tiles.

``RowExchange`` (the staged NCCL path; comparison point, and the only one
    that needs no peer mapping)
    1. all-to-all of the int32 row requests to the owners,
    2. owner-side raw row gather (``gqe_gather_rows_device``),
    3. all-to-all of the fp32 row blocks back,
    4. the same fused kernel, run against the received blocks as its tables
       with the indices rewritten to positions inside them.

``ExchangePlan`` is the host logic both share (ownership, send order, split
sizes, index rewriting); it is pure numpy and is covered by world_size-2 gloo
tests on CPU.
"""
import numpy as np

from . import _lib


def owner_by_node_type(n_modes, world):
    """mode id -> owning rank: node types are dealt round-robin over the ranks
    (8 modes on 8 GPUs = one node type per GPU, BASELINE.json configs[4])."""
    return [m % world for m in range(n_modes)]


def modes_owned_by(owner, rank):
    return [m for m, r in enumerate(owner) if r == rank]


def route_by_target_mode(formulas, mode_ids, owner):
    """Query routing for node-type-sharded tables: a formula's queries are scored on the rank
    that owns its TARGET node type.  Every query reads its positive and negative targets from
    that shard (2 of its 3-5 rows), so only the anchor rows of other node types cross NVLink:
    on the benchmark mix 35 % of the row bytes at 8 GPUs instead of 87.5 % with an arbitrary
    split -- which is what keeps the sharded path compute-bound instead of link-bound.
    -> {rank: [formula, ...]}"""
    out = {}
    for f in formulas:
        out.setdefault(owner[mode_ids[f.target_mode]], []).append(f)
    return out


class ExchangePlan(object):
    """Which of this rank's row requests go to which owner, and where the
    returned rows land.

    ``chunks`` is this rank's ordered list of operand chunks ``(mode_id, n)``:
    one per (formula segment, operand slot).  Rows are requested in the order
    (owner rank, mode id, chunk order); the owner returns them in the order it
    received them, so the reply buffer holds, per mode, ONE contiguous block
    of rows -- the staging table of that mode -- and element ``i`` of chunk
    ``c`` is row ``chunk_base[c] + i`` of its mode's staging table.
    """

    def __init__(self, owner, world, rank, chunks):
        self.owner = list(owner)
        self.world, self.rank = int(world), int(rank)
        self.n_modes = len(self.owner)
        self.chunks = [(int(m), int(n)) for m, n in chunks]
        for m, n in self.chunks:
            if not 0 <= m < self.n_modes or n < 0:
                raise ValueError("bad chunk (%d, %d)" % (m, n))
        # requests of this rank per mode, and the send order of the chunks
        self.mode_count = np.zeros(self.n_modes, dtype=np.int64)
        for m, n in self.chunks:
            self.mode_count[m] += n
        self.send_order = sorted(range(len(self.chunks)), key=lambda c: (self.owner[self.chunks[c][0]], self.chunks[c][0], c))
        # position of every chunk inside the send buffer / inside its mode's staging table
        self.chunk_send_offset = [0] * len(self.chunks)
        self.chunk_base = [0] * len(self.chunks)
        seen = np.zeros(self.n_modes, dtype=np.int64)
        pos = 0
        for c in self.send_order:
            m, n = self.chunks[c]
            self.chunk_send_offset[c] = pos
            self.chunk_base[c] = int(seen[m])
            seen[m] += n
            pos += n
        self.n_send = pos
        self.send_splits = [int(sum(self.mode_count[m] for m in range(self.n_modes) if self.owner[m] == r))
                            for r in range(self.world)]
        # first row of every mode's block inside the send / reply buffers
        self.mode_offset = np.zeros(self.n_modes, dtype=np.int64)
        pos = 0
        for r in range(self.world):
            for m in range(self.n_modes):
                if self.owner[m] == r:
                    self.mode_offset[m] = pos
                    pos += self.mode_count[m]
        self.my_modes = [m for m in range(self.n_modes) if self.owner[m] == self.rank]
        self.recv_counts = None      # [world, n_modes] after set_counts()

    def set_counts(self, counts):
        """``counts[r][m]`` = rows rank r requests of mode m (all-gathered
        ``mode_count``).  Fixes the owner-side layout."""
        counts = np.asarray(counts, dtype=np.int64).reshape(self.world, self.n_modes)
        if not np.array_equal(counts[self.rank], self.mode_count):
            raise ValueError("all-gathered counts disagree with this rank's requests")
        self.recv_counts = counts
        self.recv_splits = [int(sum(counts[s][m] for m in self.my_modes)) for s in range(self.world)]
        self.n_recv = int(sum(self.recv_splits))
        # owner side: (mode, begin, n) runs of the received request buffer
        self.recv_runs, pos = [], 0
        for s in range(self.world):
            for m in self.my_modes:
                n = int(counts[s][m])
                if n:
                    self.recv_runs.append((m, pos, n))
                pos += n
        return self

    def pack_requests(self, chunk_rows):
        """Concatenate the per-chunk row arrays (int32) in send order."""
        out = np.empty(self.n_send, dtype=np.int32)
        for c, rows in enumerate(chunk_rows):
            m, n = self.chunks[c]
            rows = np.asarray(rows).reshape(-1)
            if rows.size != n:
                raise ValueError("chunk %d has %d rows, planned %d" % (c, rows.size, n))
            out[self.chunk_send_offset[c]:self.chunk_send_offset[c] + n] = rows
        return out

    def staged_rows(self, c):
        """Rewritten indices of chunk ``c``: positions inside its mode's staging table."""
        n = self.chunks[c][1]
        return np.arange(self.chunk_base[c], self.chunk_base[c] + n, dtype=np.int32)


def chunks_of_segments(segments, n_queries_total, targets_per_query):
    """Operand chunks of a grouped batch, in a fixed order: for every segment
    its target block, then its anchor slots.  -> [(mode_id, n, kind, seg, slot)]."""
    out = []
    for si in range(len(segments)):
        sg = segments[si]
        nq = int(sg.query_end - sg.query_begin)
        out.append((int(sg.plan.target_mode), nq * targets_per_query, "target", si, 0))
        for k in range(_lib.GQE_MAX_ANCHORS):
            if sg.plan.anchor_mode[k] >= 0:
                out.append((int(sg.plan.anchor_mode[k]), nq, "anchor", si, k))
    return out


def stage_grouped(plan, chunk_info, segments, anchor_rows, target_rows, targets_per_query):
    """-> (requests int32 [n_send], staged anchor_rows, staged target_rows): the
    request vector to send to the owners and the grouped index arrays rewritten
    to address the per-mode staging tables."""
    T = targets_per_query
    chunk_rows = []
    s_anchor = np.zeros_like(anchor_rows)
    s_target = np.empty_like(target_rows)
    for c, (m, n, kind, si, k) in enumerate(chunk_info):
        qb, qe = int(segments[si].query_begin), int(segments[si].query_end)
        if kind == "target":
            chunk_rows.append(target_rows.reshape(-1)[qb * T:qe * T])
            s_target.reshape(-1)[qb * T:qe * T] = plan.staged_rows(c)
        else:
            chunk_rows.append(anchor_rows[k, qb:qe])
            s_anchor[k, qb:qe] = plan.staged_rows(c)
    return plan.pack_requests(chunk_rows), s_anchor, s_target


class RowExchange(object):
    """The staged path: request all-to-all, owner gather, row all-to-all.

    ``gather(mode_id, rows_tensor, out_tensor)`` copies raw table rows of a
    mode this rank owns; the product passes ``device_gather(ctx)`` (the CUDA
    kernel behind ``gqe_gather_rows_device``).  There is no built-in CPU
    implementation.
    """

    def __init__(self, plan, d, gather, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.plan, self.d, self.gather, self.group, self.device = plan, int(d), gather, group, device
        if plan.recv_counts is None:
            cnt = torch.from_numpy(plan.mode_count.copy())
            if device is not None:
                cnt = cnt.to(device)
            allc = [torch.empty_like(cnt) for _ in range(plan.world)]
            dist.all_gather(allc, cnt, group=group)
            plan.set_counts(torch.stack(allc).cpu().numpy())
        kw = {"device": device} if device is not None else {}
        self.req_in = torch.empty(plan.n_recv, dtype=torch.int32, **kw)
        self.rows_out = torch.empty((plan.n_recv, self.d), dtype=torch.float32, **kw)
        self.rows_in = torch.empty((plan.n_send, self.d), dtype=torch.float32, **kw)

    def run(self, requests):
        """requests: int32 tensor [n_send] in plan order -> rows tensor [n_send, d]
        (the per-mode staging tables, back to back)."""
        p, dist = self.plan, self.dist
        dist.all_to_all_single(self.req_in, requests, p.recv_splits, p.send_splits, group=self.group)
        for m, begin, n in p.recv_runs:
            self.gather(m, self.req_in[begin:begin + n], self.rows_out[begin:begin + n])
        dist.all_to_all_single(self.rows_in, self.rows_out, p.send_splits, p.recv_splits, group=self.group)
        return self.rows_in

    def staging_tables(self):
        """-> ([device pointer or 0 per mode], [rows per mode]) of the reply buffer."""
        p = self.plan
        base = self.rows_in.data_ptr()
        ptrs = [base + int(p.mode_offset[m]) * self.d * 4 if p.mode_count[m] else 0 for m in range(p.n_modes)]
        return ptrs, [int(x) for x in p.mode_count]


def device_gather(ctx):
    """Owner-side gather through the C ABI (CUDA kernel ``gqe_gather_rows``)."""
    def gather(mode_id, rows, out):
        ctx.gather_rows_device(mode_id, rows.numel(), rows.data_ptr(), out.data_ptr())
    return gather


class PeerTables(object):
    """The in-place path: map every peer's shard and hand out table pointers.

    ``local`` maps mode id -> CUDA tensor [rows, d] for the modes this rank
    owns.  ``pointers()`` returns one pointer per mode (local or peer) for
    ``Context.bind_tables``.  Collective: every rank of the group must
    construct it at the same time.
    """

    def __init__(self, ctx, owner, rows, local, group=None):
        import torch.distributed as dist
        self.ctx, self.owner, self.rows = ctx, list(owner), [int(r) for r in rows]
        rank = dist.get_rank(group)
        world = dist.get_world_size(group)
        mine = {}
        for m, own in enumerate(self.owner):
            if own == rank:
                t = local[m]
                if not t.is_cuda or not t.is_contiguous():
                    raise ValueError("shard of mode %d must be a contiguous CUDA tensor" % m)
                mine[m] = ctx.ipc_export(t.data_ptr())
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        self._local = dict(local)          # keeps the shards alive
        self._opened = []
        self._ptrs = [0] * len(self.owner)
        for m, own in enumerate(self.owner):
            if own == rank:
                self._ptrs[m] = local[m].data_ptr()
            else:
                handle, offset = everyone[own][m]
                p = ctx.ipc_open(handle, offset)
                self._opened.append(p)
                self._ptrs[m] = p
        dist.barrier(group=group)

    def pointers(self):
        return list(self._ptrs)

    def close(self):
        for p in self._opened:
            self.ctx.ipc_close(p)
        self._opened = []
