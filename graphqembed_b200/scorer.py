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


class QueryEncoderDecoder(nn.Module):
    """Encoder + metapath decoder + intersection, scored in one fused launch."""

    def __init__(self, graph, enc, path_dec, inter_dec):
        super(QueryEncoderDecoder, self).__init__()
        if not isinstance(enc, DirectEncoder):
            raise TypeError("enc must be a graphqembed_b200.DirectEncoder")
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

    # ---- context / binding ---------------------------------------------------
    def _signature(self):
        return tuple(p.data_ptr() for p in self.parameters())

    def context(self):
        p = _require_cuda(next(self.enc.parameters()), "QueryEncoderDecoder parameters")
        dev = p.device.index if p.device.index is not None else torch.cuda.current_device()
        st = self._state
        if st is None or st[1] != dev:
            st = [_lib.Context(dev), dev, None]
            self._state = st
        sig = self._signature()
        if st[2] != sig:
            self.enc._bind(st[0])
            self.path_dec._bind(st[0])
            if isinstance(self.inter_dec, SetIntersection):
                modes = self.enc.modes
                pre = [_require_cuda(self.inter_dec.pre_mats[m], "pre matrix").data_ptr() for m in modes]
                post = [_require_cuda(self.inter_dec.post_mats[m], "post matrix").data_ptr() for m in modes]
                st[0].bind_intersection(_lib.INTER_ID[self.inter_dec.kind], pre, post, self.inter_dec.dim,
                                        self.inter_dec.expand_dim)
            else:
                st[0].bind_intersection(_lib.INTER_ID[self.inter_dec.kind], None, None, self.enc.dim)
            st[2] = sig
        st[0].set_stream(torch.cuda.current_stream(dev).cuda_stream)
        st[0].set_precision(self.precision)
        st[0].set_compose(self.compose)
        return st[0]

    @property
    def device(self):
        return next(self.enc.parameters()).device

    def plan(self, formula):
        """Lowered formula (cached): structure id, mode ids, relation ids."""
        pl = self._plans.get(formula)
        if pl is None:
            pl = lower_formula(formula, self.enc.mode_ids, self.path_dec.rel_ids)
            self._plans[formula] = pl
        return pl

    # ---- index lowering --------------------------------------------------------
    def lower_batch(self, batch):
        """QueryBatch (node ids) -> (anchor_rows [A,Q] int32, target_rows [P] int32)."""
        f = batch.formula
        nq = batch.n_queries
        anchor_rows = np.empty((len(f.anchor_modes), nq), dtype=np.int32)
        for k, mode in enumerate(f.anchor_modes):
            anchor_rows[k] = self.enc.rows(batch.anchors[k], mode)
        target_rows = self.enc.rows(batch.targets, f.target_mode)
        return anchor_rows, target_rows

    def _to_dev(self, arr):
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.device, non_blocking=True)

    # ---- scoring -----------------------------------------------------------------
    def score_batch(self, batch):
        """Scores of a QueryBatch, in the batch's own pair order -> FloatTensor[n_pairs]."""
        ctx = self.context()
        plan = self.plan(batch.formula)
        anchor_rows, target_rows = self.lower_batch(batch)
        a = self._to_dev(anchor_rows)
        t = self._to_dev(target_rows)
        off = None if batch.offsets is None else self._to_dev(batch.offsets)
        out = torch.empty(batch.n_pairs, dtype=torch.float32, device=self.device)
        ctx.score_device(plan, batch.n_queries, a.data_ptr(), batch.n_pairs, t.data_ptr(),
                         None if off is None else off.data_ptr(), out.data_ptr())
        return out

    def forward(self, formula, queries, source_nodes):
        """model.py:70-109.  Unknown query types return None like the reference."""
        if formula.query_type not in QUERY_TYPES:
            return None
        batch, order = QueryBatch.from_queries(formula, queries, source_nodes)
        scores = self.score_batch(batch)
        if order is None:
            return scores
        out = torch.empty_like(scores)
        out[self._to_dev(order)] = scores
        return out

    def pick_negatives(self, formula, queries, hard_negatives=False):
        """model.py:113-120 -- same draws from the global ``random`` stream."""
        if "inter" not in formula.query_type and hard_negatives:
            raise Exception("Hard negative examples can only be used with intersection queries")
        if hard_negatives:
            return [random.choice(query.hard_neg_samples) for query in queries]
        if formula.query_type == "1-chain":
            return [random.choice(self.graph.full_lists[formula.target_mode]) for _ in queries]
        return [random.choice(query.neg_samples) for query in queries]

    def margin_loss(self, formula, queries, hard_negatives=False, margin=1):
        """model.py:112-127 in one launch: the query side is built once and
        scored against the positive and the negative; hinge + mean are fused."""
        neg_nodes = self.pick_negatives(formula, queries, hard_negatives)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training: the differentiable operator chain (autograd.py); its backward runs the
            # hand-written VJP kernels and leaves dense .grad tensors for any torch optimiser
            from . import autograd
            return autograd.margin_loss(self, formula, queries, neg_nodes, margin)
        n = len(queries)
        anchors = np.empty((len(formula.anchor_modes), n), dtype=np.int64)
        for k in range(anchors.shape[0]):
            anchors[k] = np.fromiter((q.anchor_nodes[k] for q in queries), dtype=np.int64, count=n)
        pairs = np.empty((n, 2), dtype=np.int64)
        pairs[:, 0] = np.fromiter((q.target_node for q in queries), dtype=np.int64, count=n)
        pairs[:, 1] = np.fromiter(neg_nodes, dtype=np.int64, count=n)
        return self.margin_loss_batch(QueryBatch(formula, anchors, pairs.reshape(-1)), margin)

    def margin_loss_batch(self, batch, margin=1, return_scores=False):
        """Loss of a QueryBatch whose targets are (positive, negative) per query."""
        if batch.offsets is not None or batch.n_pairs != 2 * batch.n_queries:
            raise ValueError("margin loss needs exactly (positive, negative) per query")
        ctx = self.context()
        plan = self.plan(batch.formula)
        anchor_rows, pair_rows = self.lower_batch(batch)
        a = self._to_dev(anchor_rows)
        t = self._to_dev(pair_rows)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        scores = torch.empty((batch.n_queries, 2), dtype=torch.float32, device=self.device) if return_scores else None
        ctx.margin_loss_device(plan, batch.n_queries, a.data_ptr(), t.data_ptr(), margin, loss.data_ptr(),
                               None if scores is None else scores.data_ptr())
        return (loss, scores) if return_scores else loss

    def margin_loss_grouped(self, batches, margin=1, return_scores=False):
        """One call for many formulas (the "full mix" workload): mean hinge over
        ALL queries of all batches.  Each batch holds (positive, negative) pairs."""
        ctx = self.context()
        total = sum(b.n_queries for b in batches)
        anchor_rows = np.zeros((_lib.GQE_MAX_ANCHORS, total), dtype=np.int32)
        pair_rows = np.empty((total, 2), dtype=np.int32)
        items, q0 = [], 0
        for b in batches:
            if b.offsets is not None or b.n_pairs != 2 * b.n_queries:
                raise ValueError("margin loss needs exactly (positive, negative) per query")
            a, t = self.lower_batch(b)
            anchor_rows[:a.shape[0], q0:q0 + b.n_queries] = a
            pair_rows[q0:q0 + b.n_queries] = t.reshape(-1, 2)
            items.append((self.plan(b.formula), q0, q0 + b.n_queries))
            q0 += b.n_queries
        segs = _lib.make_segments(items)
        a = self._to_dev(anchor_rows)
        t = self._to_dev(pair_rows)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        scores = torch.empty((total, 2), dtype=torch.float32, device=self.device) if return_scores else None
        ctx.score_grouped_device(segs, total, a.data_ptr(), t.data_ptr(), 2,
                                 None if scores is None else scores.data_ptr(), margin, loss.data_ptr())
        return (loss, scores) if return_scores else loss
