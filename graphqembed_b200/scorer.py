"""``QueryEncoderDecoder``: the fused conjunctive-query scorer.

Drop-in for reference ``netquery/model.py:57-127``: same constructor
``(graph, enc, path_dec, inter_dec)``, same ``forward(formula, queries,
source_nodes) -> FloatTensor[B]`` and ``margin_loss(formula, queries,
hard_negatives=False, margin=1) -> 0-dim tensor``.  Where the reference runs
~20 ATen ops per call, this issues ONE kernel per formula: embedding gather,
L2 normalisation, the chained relation operators, the intersection, the
cosine against every target and (for the loss) the hinge + mean all happen on
chip (``csrc/gqe_simt.cuh``).

``forward`` and the ``*_batch`` / ``*_grouped`` calls never build an autograd
graph.  ``margin_loss`` does when gradients are enabled and a parameter
requires them (the training loop of reference train_helpers.py:76-79): it then
runs the differentiable operator chain of ``autograd.py`` (exact fp32, un-fused)
so that ``loss.backward(); optimizer.step()`` works unchanged; under
``torch.no_grad()`` it is the single fused launch.
"""
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .lowering import lower_formula
from .operators import DirectEncoder, SetIntersection, SimpleSetIntersection, _MetapathDecoder, _require_cuda
from .query import QUERY_TYPES, QueryBatch
from .store import DeviceSlice, StoreSlice


class QueryEncoderDecoder(nn.Module):
    """Encoder + metapath decoder + intersection, scored in one fused launch."""

    def __init__(self, graph, enc, path_dec, inter_dec):
        super(QueryEncoderDecoder, self).__init__()
        from .sage import Encoder as SageEncoder
        if not isinstance(enc, (DirectEncoder, SageEncoder)):
            raise TypeError("enc must be a graphqembed_b200.DirectEncoder or a graphqembed_b200.sage.Encoder")
        # the fused kernels gather + normalise table rows themselves: DirectEncoder only.  With a
        # GraphSAGE-style encoder the operators run one after another as in model.py:70-109
        self._fused = isinstance(enc, DirectEncoder)
        if not isinstance(path_dec, _MetapathDecoder):
            raise TypeError("path_dec must be a graphqembed_b200 metapath decoder")
        if not isinstance(inter_dec, (SetIntersection, SimpleSetIntersection)):
            raise TypeError("inter_dec must be a graphqembed_b200 intersection operator")
        self.enc = enc
        self.path_dec = path_dec
        self.inter_dec = inter_dec
        self.graph = graph
        if enc.dim != path_dec.dim or (isinstance(inter_dec, SetIntersection) and inter_dec.dim != enc.dim):
            raise ValueError("encoder, decoder and intersection dimensions differ")
        self._plans = {}
        self._state = None   # [Context, device index, pointer signature]
        # arithmetic of the d x d contractions: "bf16x3" = tcgen05 tensor cores with
        # split-bf16 operands (default), "fp32" = exact CUDA-core FMA (include/gqe.h)
        self.precision = "bf16x3"
        # tensor-core path only: pre-multiply runs of linear operators once per call
        # ("auto": when >= 1024 rows of a formula share the product; "off"; "always")
        self.compose = "auto"
        # raise KeyError / IndexError for a bad node id inside the call, like the reference (one
        # stream synchronisation per call); off = asynchronous, poll context().index_error()
        self.check_indices = True
        # negatives of StoreSlice batches: a numpy Generator (None: a process-wide default), or
        # the reference's own global-``random`` draws when reference_negatives is set
        self.negative_rng = None
        self.reference_negatives = False
        # negatives of DeviceSlice batches are drawn on the GPU by a counter-based generator: the seed of call k
        # is negative_seed + k (None: a random base taken once per model)
        self.negative_seed = None
        self._store_calls = 0
        self._store_cache = {}  # descriptors of recent DeviceSlice calls: (key) -> (segments, slices, n, keep-alive)
        # training with optim.SparseRowAdam: table gradients as (row, gradient) pairs and a hook
        # called with the rows a differentiable forward is about to gather
        self.sparse_table_grads = False
        self.row_hook = None
        self._stage = None     # [pinned buffer, device buffer, numpy view of the pinned one, last copy's event]
        self._cache = None     # (parameters, operator parameters, device)
        self._knobs = None

    # ---- context / binding ---------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() may move or replace parameter storage: re-bind on the next call
        out = super(QueryEncoderDecoder, self)._apply(fn, *args, **kwargs)
        self._cache = None
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super(QueryEncoderDecoder, self).load_state_dict(*args, **kwargs)
        self._cache = None
        return out

    def _lists(self):
        """(all parameters, operator matrices / vectors, device) -- cached: walking the module tree costs
        more than a small scoring call."""
        c = self._cache
        if c is None:
            allp = list(self.parameters())
            ops = list(self.path_dec.parameters()) + list(self.inter_dec.parameters())
            p0 = _require_cuda(next(self.enc.parameters()), "QueryEncoderDecoder parameters")
            c = self._cache = (allp, ops, p0.device)
        return c

    def _signature(self):
        return tuple(p.data_ptr() for p in self._lists()[0])

    def _weight_version(self):
        """Sum of the autograd version counters of every operator matrix / vector: changes
        whenever one of them is modified in place (optimizer.step(), copy_, load_state_dict)."""
        v = 0
        for p in self._lists()[1]:
            v += p._version
        return v

    def context(self):
        if not self._fused:
            raise RuntimeError("the fused calls gather table rows themselves and need a DirectEncoder; with a "
                               "GraphSAGE-style encoder use forward() / margin_loss() (un-fused operator chain)")
        allp, _, device = self._lists()
        dev = device.index if device.index is not None else torch.cuda.current_device()
        st = self._state
        if st is None or st[1] != dev:
            st = [_lib.Context(dev), dev, None, None, False, None]   # ctx, device, pointers, weight version, node maps?, keep-alive
            self._state = st
        ctx = st[0]
        sig = self._signature()
        if st[2] != sig:
            self.enc._bind(ctx)
            self.path_dec._bind(ctx)
            if isinstance(self.inter_dec, SetIntersection):
                modes = self.enc.modes
                pre = [_require_cuda(self.inter_dec.pre_mats[m], "pre matrix").data_ptr() for m in modes]
                post = [_require_cuda(self.inter_dec.post_mats[m], "post matrix").data_ptr() for m in modes]
                ctx.bind_intersection(_lib.INTER_ID[self.inter_dec.kind], pre, post, self.inter_dec.dim,
                                      self.inter_dec.expand_dim)
            else:
                ctx.bind_intersection(_lib.INTER_ID[self.inter_dec.kind], None, None, self.enc.dim)
            # the node id -> row tables of the modes, uploaded once: the kernels then take node ids
            maps = self.enc.features.device_maps(self.enc.modes, [self.enc.table(m).size(0) for m in self.enc.modes],
                                                 device)
            st[4] = maps is not None
            if maps is not None:
                ctx.bind_node_maps(maps[0], maps[1], maps[2])
                st[5] = maps[3]
            st[2] = sig
            st[3] = None
        # packed / pre-multiplied operator matrices are cached inside the context; an in-place
        # update of any of them (autograd version counter) drops the cache
        ver = self._weight_version()
        if st[3] != ver:
            ctx.invalidate_weights()
            st[3] = ver
        stream = torch.cuda.current_stream(dev).cuda_stream
        knobs = (stream, self.precision, self.compose)
        if self._knobs != knobs:
            ctx.set_stream(stream)
            ctx.set_precision(self.precision)
            ctx.set_compose(self.compose)
            self._knobs = knobs
        return ctx

    def __getstate__(self):
        # the native context (a ctypes handle) is per process: recreated lazily after
        # copy.deepcopy / pickle / torch.save(model)
        state = dict(self.__dict__)
        state["_state"] = None
        state["_stage"] = None
        state["_cache"] = None
        state["_knobs"] = None
        return state

    @property
    def device(self):
        return self._lists()[2]

    def plan(self, formula):
        """Lowered formula (cached): structure id, mode ids, relation ids."""
        pl = self._plans.get(formula)
        if pl is None:
            pl = lower_formula(formula, self.enc.mode_ids, self.path_dec.rel_ids)
            self._plans[formula] = pl
        return pl

    # ---- index lowering --------------------------------------------------------
    def lower_batch(self, batch):
        """QueryBatch (node ids) -> (anchor_rows [A,Q] int32, target_rows [P] int32) on the host
        (an O(1) table lookup per node; KeyError on an unknown node)."""
        f = batch.formula
        nq = batch.n_queries
        anchor_rows = np.empty((len(f.anchor_modes), nq), dtype=np.int32)
        for k, mode in enumerate(f.anchor_modes):
            anchor_rows[k] = self.enc.rows(batch.anchors[k], mode)
        target_rows = self.enc.rows(batch.targets, f.target_mode)
        return anchor_rows, target_rows

    def _indices(self, batch):
        """-> (anchors int32 [A,Q], targets int32 [P], nodes): the node ids themselves when the
        context holds the node maps (the kernels do the lookup of bio/data_utils.py:20-21), else
        rows lowered on the host."""
        if self._state is not None and self._state[4]:
            ids = batch.int32_ids()
            if ids is not None:
                return ids[0], ids[1], True
        a, t = self.lower_batch(batch)
        return a, t, False

    def _to_dev(self, *arrays):
        """numpy arrays -> device tensors through ONE reusable pinned staging buffer and ONE
        asynchronous H2D copy (torch's own pageable path is a synchronous staging copy per array
        inside the driver)."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        offs, total = [], 0
        for a in arrays:
            offs.append(total)
            total += (a.nbytes + 255) & ~255
        st = self._stage
        if st is None or st[0].numel() < total or st[1].device != self.device:
            cap = max(total + total // 2, 1 << 16)
            pin = torch.empty(cap, dtype=torch.uint8).pin_memory()
            st = self._stage = [pin, torch.empty(cap, dtype=torch.uint8, device=self.device), pin.numpy(), None]
        if st[3] is not None:
            st[3].synchronize()              # the previous copy out of the pinned buffer has landed
        out = []
        for a, o in zip(arrays, offs):
            st[2][o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
            out.append(st[1][o:o + a.nbytes].view(torch.from_numpy(a[:0]).dtype).view(a.shape))
        if total:
            st[1][:total].copy_(st[0][:total], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            st[3] = ev
        return out if len(out) > 1 else out[0]

    def _check_indices(self, ctx, nodes):
        """The reference raises KeyError inside forward() for a node that is not in node_maps;
        with the lookup on the device the kernels report it asynchronously.  ``check_indices``
        (default on) synchronises and raises here, like the reference; turn it off to keep the
        call asynchronous and poll ``context().index_error()`` yourself."""
        if self.check_indices:
            ctx.index_error()

    # ---- scoring -----------------------------------------------------------------
    def score_batch(self, batch):
        """Scores of a QueryBatch, in the batch's own pair order -> FloatTensor[n_pairs]."""
        ctx = self.context()
        plan = self.plan(batch.formula)
        anchors, targets, nodes = self._indices(batch)
        if batch.offsets is None:
            (a, t), off = self._to_dev(anchors, targets), None
        else:
            a, t, off = self._to_dev(anchors, targets, batch.offsets)
        out = torch.empty(batch.n_pairs, dtype=torch.float32, device=self.device)
        ctx.score_device(plan, batch.n_queries, a.data_ptr(), batch.n_pairs, t.data_ptr(),
                         None if off is None else off.data_ptr(), out.data_ptr(), nodes=nodes)
        self._check_indices(ctx, nodes)
        return out

    def forward(self, formula, queries, source_nodes):
        """model.py:70-109.  Unknown query types return None like the reference."""
        if formula.query_type not in QUERY_TYPES:
            return None
        if not self._fused:
            return self._forward_unfused(formula, queries, source_nodes)
        if isinstance(queries, StoreSlice):      # pair i = queries[i] against source_nodes[i]
            return self.score_batch(QueryBatch(formula, queries.anchors, np.asarray(source_nodes, dtype=np.int32)))
        batch, order = QueryBatch.from_queries(formula, queries, source_nodes)
        scores = self.score_batch(batch)
        if order is None:
            return scores
        out = torch.empty_like(scores)
        out[self._to_dev(order)] = scores
        return out

    def _forward_unfused(self, formula, queries, source_nodes):
        """model.py:70-109 operator by operator (encoder calls in the reference's order: the
        GraphSAGE-style encoder draws its neighbour samples from the global ``random`` stream)."""
        from .operators import cosine_similarity_dim0
        from .query import reverse_relation as rev
        enc, dec, qt = self.enc, self.path_dec, formula.query_type
        anchors = lambda k: [q.anchor_nodes[k] for q in queries]
        if qt in ("1-chain", "2-chain", "3-chain"):
            return dec.forward(enc.forward(source_nodes, formula.target_mode),
                               enc.forward(anchors(0), formula.anchor_modes[0]), formula.rels)
        target = enc.forward(source_nodes, formula.target_mode)
        if qt == "3-chain_inter":
            e1 = dec.project(enc.forward(anchors(0), formula.anchor_modes[0]), rev(formula.rels[1][0]))
            e2 = dec.project(enc.forward(anchors(1), formula.anchor_modes[1]), rev(formula.rels[1][1]))
            q = dec.project(self.inter_dec(e1, e2, formula.rels[0][-1]), rev(formula.rels[0]))
            return cosine_similarity_dim0(target, q)
        e1 = dec.project(enc.forward(anchors(0), formula.anchor_modes[0]), rev(formula.rels[0]))
        e2 = enc.forward(anchors(1), formula.anchor_modes[1])
        if len(formula.rels[1]) == 2:
            for i_rel in formula.rels[1][::-1]:
                e2 = dec.project(e2, rev(i_rel))
        else:
            e2 = dec.project(e2, rev(formula.rels[1]))
        if qt == "3-inter":
            e3 = dec.project(enc.forward(anchors(2), formula.anchor_modes[2]), rev(formula.rels[2]))
            q = self.inter_dec(e1, e2, formula.target_mode, e3)
        else:
            q = self.inter_dec(e1, e2, formula.target_mode)
        return cosine_similarity_dim0(target, q)

    def pick_negatives(self, formula, queries, hard_negatives=False):
        """model.py:113-120 -- same draws from the global ``random`` stream."""
        if "inter" not in formula.query_type and hard_negatives:
            raise Exception("Hard negative examples can only be used with intersection queries")
        if hard_negatives:
            return [random.choice(query.hard_neg_samples) for query in queries]
        if formula.query_type == "1-chain":
            return [random.choice(self.graph.full_lists[formula.target_mode]) for _ in queries]
        return [random.choice(query.neg_samples) for query in queries]

    def _full_array(self, mode):
        g = self.graph
        if hasattr(g, "full_array"):
            return g.full_array(mode)
        cache = self.__dict__.setdefault("_full_arrays", {})
        arr = cache.get(mode)
        if arr is None:
            arr = cache[mode] = np.asarray(g.full_lists[mode], dtype=np.int32)
        return arr

    def margin_loss(self, formula, queries, hard_negatives=False, margin=1):
        """model.py:112-127 in one launch: the query side is built once and
        scored against the positive and the negative; hinge + mean are fused.

        ``queries``: a list of ``Query`` objects (the reference's argument), or a
        ``StoreSlice`` of a ``QueryStore`` -- then nothing here is per-query Python: the
        anchors / targets are array views, the negatives one vectorised draw
        (``self.negative_rng``; ``self.reference_negatives = True`` draws with the global
        ``random`` module in the reference's order instead)."""
        if not self._fused:
            # inference only: two operator-chain passes like model.py:122-127
            neg_nodes = self.pick_negatives(formula, queries, hard_negatives)
            affs = self.forward(formula, queries, [q.target_node for q in queries])
            neg_affs = self.forward(formula, queries, neg_nodes)
            return torch.clamp(margin - (affs - neg_affs), min=0).mean()
        if isinstance(queries, DeviceSlice):
            if self._store_on_device():
                return self._margin_loss_store([(formula, queries)], hard_negatives, margin)
            queries = queries.host()        # training / no node maps on the device: the host arrays of the same slice
        if isinstance(queries, StoreSlice):
            if "inter" not in formula.query_type and hard_negatives:
                raise Exception("Hard negative examples can only be used with intersection queries")
            full = self._full_array(formula.target_mode) if formula.query_type == "1-chain" else None
            neg_nodes = queries.draw_negatives(hard_negatives, full, self.negative_rng, self.reference_negatives)
            anchors, pos_nodes = queries.anchors, queries.targets
        else:
            neg_nodes = self.pick_negatives(formula, queries, hard_negatives)
            n = len(queries)
            anchors = np.empty((len(formula.anchor_modes), n), dtype=np.int64)
            for k in range(anchors.shape[0]):
                anchors[k] = np.fromiter((q.anchor_nodes[k] for q in queries), dtype=np.int64, count=n)
            pos_nodes = np.fromiter((q.target_node for q in queries), dtype=np.int64, count=n)
            neg_nodes = np.fromiter(neg_nodes, dtype=np.int64, count=n)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._lists()[0]):
            # training: the differentiable operator chain (autograd.py); its backward runs the
            # hand-written VJP kernels and leaves .grad tensors for the optimiser
            from . import autograd
            return autograd.margin_loss(self, formula, list(anchors), pos_nodes, neg_nodes, margin)
        pairs = np.empty((len(pos_nodes), 2), dtype=anchors.dtype)
        pairs[:, 0] = pos_nodes
        pairs[:, 1] = neg_nodes
        return self.margin_loss_batch(QueryBatch(formula, anchors, pairs.reshape(-1)), margin)

    def margin_loss_batch(self, batch, margin=1, return_scores=False):
        """Loss of a QueryBatch whose targets are (positive, negative) per query."""
        if batch.offsets is not None or batch.n_pairs != 2 * batch.n_queries:
            raise ValueError("margin loss needs exactly (positive, negative) per query")
        ctx = self.context()
        plan = self.plan(batch.formula)
        anchors, pairs, nodes = self._indices(batch)
        a, t = self._to_dev(anchors, pairs)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        scores = torch.empty((batch.n_queries, 2), dtype=torch.float32, device=self.device) if return_scores else None
        ctx.margin_loss_device(plan, batch.n_queries, a.data_ptr(), t.data_ptr(), margin, loss.data_ptr(),
                               None if scores is None else scores.data_ptr(), nodes=nodes)
        self._check_indices(ctx, nodes)
        return (loss, scores) if return_scores else loss

    def margin_loss_mix(self, items, hard_negatives=False, margin=1):
        """The margin loss of MANY formulas' batches in one launch: ``items`` = [(formula, StoreSlice)].
        Negatives are drawn per slice exactly as ``margin_loss(formula, slice)`` draws them; the result is
        the mean hinge over all queries of all slices (= the size-weighted mean of the per-formula
        losses).  What a loop over ``margin_loss`` costs per formula -- a host call, an H2D copy, a
        small kernel, a loss read -- is paid once."""
        if items and all(isinstance(sl, DeviceSlice) for _, sl in items):
            if self._store_on_device():
                return self._margin_loss_store(items, hard_negatives, margin)
            items = [(f, sl.host()) for f, sl in items]
        batches = []
        for formula, sl in items:
            if "inter" not in formula.query_type and hard_negatives:
                raise Exception("Hard negative examples can only be used with intersection queries")
            full = self._full_array(formula.target_mode) if formula.query_type == "1-chain" else None
            batches.append(sl.margin_batch(sl.draw_negatives(hard_negatives, full, self.negative_rng, self.reference_negatives)))
        return self.margin_loss_grouped(batches, margin)

    # ---- device-resident store ---------------------------------------------------------
    def _store_on_device(self):
        """DeviceSlice batches run natively when this is a scoring call (no autograd graph wanted)
        and the context holds the node maps (the store holds node ids)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._lists()[0]):
            return False
        if self._state is None:
            self.context()
        return bool(self._state[4])

    def _full_device(self, mode):
        cache = self.__dict__.setdefault("_full_dev", {})
        t = cache.get(mode)
        if t is None or t.device != self.device:
            t = cache[mode] = torch.from_numpy(np.ascontiguousarray(self._full_array(mode), dtype=np.int32)).to(self.device)
        return t

    def _margin_loss_store(self, items, hard_negatives=False, margin=1, return_pairs=False, return_scores=False):
        """[(formula, DeviceSlice)] -> loss (0-d device tensor) through gqe_margin_loss_store_device:
        slices gathered and negatives drawn by one small kernel, the fused scoring kernel behind it."""
        ctx = self.context()
        key = (hard_negatives,) + tuple((f, id(sl.block), sl.start, sl.stop) for f, sl in items)
        ent = self._store_cache.get(key)
        if ent is None:
            seg_items, q0, keep = [], 0, []
            arr = (_lib.StoreSliceC * len(items))()
            for i, (formula, sl) in enumerate(items):
                if "inter" not in formula.query_type and hard_negatives:
                    raise Exception("Hard negative examples can only be used with intersection queries")
                if sl.block.anchors.device != self.device:
                    raise ValueError("the store lives on %s, the model on %s" % (sl.block.anchors.device, self.device))
                b = sl.block
                arr[i].anchors, arr[i].targets = b.anchors.data_ptr(), b.targets.data_ptr()
                arr[i].block_queries, arr[i].start, arr[i].pool_size = b.n, sl.start, 0
                if hard_negatives:
                    arr[i].neg_ptr, arr[i].negs = b.hard_ptr.data_ptr(), b.hards.data_ptr()
                elif formula.query_type == "1-chain":     # any node of the target mode (model.py:118-119)
                    pool = self._full_device(formula.target_mode)
                    keep.append(pool)
                    arr[i].neg_ptr, arr[i].negs, arr[i].pool_size = None, pool.data_ptr(), pool.numel()
                else:
                    arr[i].neg_ptr, arr[i].negs = b.neg_ptr.data_ptr(), b.negs.data_ptr()
                seg_items.append((self.plan(formula), q0, q0 + len(sl)))
                q0 += len(sl)
                keep.append(b)
            if len(self._store_cache) >= 64:
                self._store_cache.clear()
            ent = self._store_cache[key] = (_lib.make_segments(seg_items), arr, q0, keep)
        segs, arr, nq, _ = ent
        if self.negative_seed is None:
            self.negative_seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0]) >> 1
        seed = self.negative_seed + self._store_calls
        self._store_calls += 1
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        pairs = torch.empty((nq, 2), dtype=torch.int32, device=self.device) if return_pairs else None
        scores = torch.empty((nq, 2), dtype=torch.float32, device=self.device) if return_scores else None
        ctx.margin_loss_store_device(segs, arr, seed, margin, loss.data_ptr(),
                                     None if scores is None else scores.data_ptr(),
                                     None if pairs is None else pairs.data_ptr())
        self._check_indices(ctx, True)
        if return_pairs or return_scores:
            return loss, pairs, scores
        return loss

    def margin_loss_grouped(self, batches, margin=1, return_scores=False):
        """One call for many formulas (the "full mix" workload): mean hinge over
        ALL queries of all batches.  Each batch holds (positive, negative) pairs."""
        ctx = self.context()
        total = sum(b.n_queries for b in batches)
        anchor_idx = np.zeros((_lib.GQE_MAX_ANCHORS, total), dtype=np.int32)
        pair_idx = np.empty((total, 2), dtype=np.int32)
        lowered = [self._indices(b) for b in batches]
        nodes = all(x[2] for x in lowered)
        items, q0 = [], 0
        for b, (a, t, was_nodes) in zip(batches, lowered):
            if b.offsets is not None or b.n_pairs != 2 * b.n_queries:
                raise ValueError("margin loss needs exactly (positive, negative) per query")
            if was_nodes and not nodes:          # mixed: fall back to host-lowered rows for all
                a, t = self.lower_batch(b)
            anchor_idx[:a.shape[0], q0:q0 + b.n_queries] = a
            pair_idx[q0:q0 + b.n_queries] = t.reshape(-1, 2)
            items.append((self.plan(b.formula), q0, q0 + b.n_queries))
            q0 += b.n_queries
        segs = _lib.make_segments(items)
        a, t = self._to_dev(anchor_idx, pair_idx)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        scores = torch.empty((total, 2), dtype=torch.float32, device=self.device) if return_scores else None
        ctx.score_grouped_device(segs, total, a.data_ptr(), t.data_ptr(), 2,
                                 None if scores is None else scores.data_ptr(), margin, loss.data_ptr(), nodes=nodes)
        self._check_indices(ctx, nodes)
        return (loss, scores) if return_scores else loss
