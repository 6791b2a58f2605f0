"""Differentiable scoring path: ``loss.backward()`` on the GPU.

The reference trains by back-propagating ``margin_loss`` through its PyTorch
ops (``netquery/train_helpers.py:76-79``).  Here every operator of the path is
a ``torch.autograd.Function`` whose forward AND backward are the hand-written
kernels behind the C ABI (``csrc/gqe_simt.cuh`` forward pieces,
``csrc/gqe_bwd.cu`` vector-Jacobian products); autograd only chains them, the
way it chains the reference's ops:

=====================  ==========================  ================================
reference op           forward kernel              backward kernel
=====================  ==========================  ================================
encoders.py:41-43      gqe_encode_device           gqe_encode_bwd_device (dense
                                                   table gradient, like nn.Embedding)
decoders.py:145,150,   gqe_matmul_device           gqe_matmul_device (transposed) +
289,299 (mm)                                       gqe_matmul_wgrad_device
decoders.py:203,208,   gqe_project_device          identity / gqe_project_device +
231,236 (vectors)                                  gqe_rowsum_device
decoders.py:289-296,   gqe_aggregate_device        gqe_aggregate_bwd_device
311-316 (relu, agg)
model.py:68 (cosine),  gqe_cosine_device /         gqe_cosine_bwd_device
decoders.py:232 (dot)  gqe_dot_device
=====================  ==========================  ================================

The hinge + mean over the B scores (model.py:124-126) are plain torch ops on a
length-B vector.  This path is exact fp32 and un-fused (every intermediate
``[d, B]`` tensor lives in HBM); the fused kernels serve ``forward`` /
``margin_loss`` whenever no gradient is required.
"""
import numpy as np
import torch
from torch.autograd import Function

from .query import CHAIN_TYPES, FLAT_INTER_TYPES, reverse_relation


def _c(t):
    return t.contiguous() if not t.is_contiguous() else t


class _Encode(Function):
    """DirectEncoder.forward (encoders.py:41-43).  Backward: the dense [rows, d] table gradient
    nn.Embedding produces (``sparse`` False), or the same gradient as a sparse COO tensor of
    (gathered row, gradient row) pairs -- no dense tensor, no atomics -- for SparseRowAdam."""

    @staticmethod
    def forward(fctx, weight, ctx, mode_id, rows, sparse):
        d, n = weight.size(1), rows.numel()
        out = torch.empty((d, n), dtype=torch.float32, device=weight.device)
        ctx.encode_device(mode_id, n, rows.data_ptr(), out.data_ptr())
        fctx.gqe = (ctx, mode_id, rows, tuple(weight.shape), sparse)
        return out

    @staticmethod
    def backward(fctx, g):
        ctx, mode_id, rows, shape, sparse = fctx.gqe
        g = _c(g)
        if sparse:
            vals = torch.empty((rows.numel(), shape[1]), dtype=torch.float32, device=g.device)
            ctx.encode_bwd_rows_device(mode_id, rows.numel(), rows.data_ptr(), g.data_ptr(), vals.data_ptr())
            return torch.sparse_coo_tensor(rows.to(torch.int64).unsqueeze(0), vals, shape), None, None, None, None
        gtable = torch.zeros(shape, dtype=torch.float32, device=g.device)
        ctx.encode_bwd_device(mode_id, rows.numel(), rows.data_ptr(), g.data_ptr(), gtable.data_ptr())
        return gtable, None, None, None, None


class _Matmul(Function):
    """y = W x (transpose False: M.mm(embeds)) or W^T x (transpose True: act.mm(M))."""

    @staticmethod
    def forward(fctx, w, x, ctx, transpose):
        x = _c(x)
        out = torch.empty_like(x)
        ctx.matmul_device(w.data_ptr(), transpose, x.size(0), x.size(1), x.data_ptr(), out.data_ptr())
        fctx.save_for_backward(w, x)
        fctx.gqe = (ctx, transpose)
        return out

    @staticmethod
    def backward(fctx, g):
        w, x = fctx.saved_tensors
        ctx, transpose = fctx.gqe
        g = _c(g)
        d, n = x.size(0), x.size(1)
        gx = gw = None
        if fctx.needs_input_grad[1]:
            gx = torch.empty_like(x)
            ctx.matmul_device(w.data_ptr(), not transpose, d, n, g.data_ptr(), gx.data_ptr())
        if fctx.needs_input_grad[0]:
            gw = torch.zeros_like(w)
            ctx.matmul_wgrad_device(transpose, d, n, g.data_ptr(), x.data_ptr(), gw.data_ptr())
        return gw, gx, None, None


class _VecOp(Function):
    """y = x + v (TransE, decoders.py:203,208) or x * v (BilinearDiag, decoders.py:231,236);
    ``rid`` is the relation id whose vector ``v`` is bound in ``ctx``."""

    @staticmethod
    def forward(fctx, v, x, ctx, rid, mul):
        x = _c(x)
        out = torch.empty_like(x)
        ctx.project_device(rid, x.size(1), x.data_ptr(), out.data_ptr())
        fctx.save_for_backward(x)
        fctx.gqe = (ctx, rid, mul, v.shape)
        return out

    @staticmethod
    def backward(fctx, g):
        (x,) = fctx.saved_tensors
        ctx, rid, mul, vshape = fctx.gqe
        g = _c(g)
        d, n = x.size(0), x.size(1)
        gv = torch.zeros(vshape, dtype=torch.float32, device=g.device)
        ctx.rowsum_device(d, n, g.data_ptr(), x.data_ptr() if mul else None, gv.data_ptr())
        if mul:
            gx = torch.empty_like(x)
            ctx.project_device(rid, n, g.data_ptr(), gx.data_ptr())      # g * v
        else:
            gx = g
        return gv, gx, None, None, None


class _Aggregate(Function):
    """agg_k act(e_k): act = relu | identity, agg = mean | min (decoders.py:289-296,311-316)."""

    @staticmethod
    def forward(fctx, ctx, relu, use_min, e1, e2, e3):
        e1, e2 = _c(e1), _c(e2)
        e3 = None if e3 is None else _c(e3)
        out = torch.empty_like(e1)
        ctx.aggregate_device(e1.size(0), e1.size(1), e1.data_ptr(), e2.data_ptr(), None if e3 is None else e3.data_ptr(),
                             relu, use_min, out.data_ptr())
        fctx.save_for_backward(*([e1, e2] + ([] if e3 is None else [e3])))
        fctx.gqe = (ctx, relu, use_min)
        return out

    @staticmethod
    def backward(fctx, g):
        es = fctx.saved_tensors
        ctx, relu, use_min = fctx.gqe
        g = _c(g)
        gs = [torch.empty_like(e) for e in es]
        e3 = es[2].data_ptr() if len(es) > 2 else None
        g3 = gs[2].data_ptr() if len(es) > 2 else None
        ctx.aggregate_bwd_device(es[0].size(0), es[0].size(1), es[0].data_ptr(), es[1].data_ptr(), e3, relu, use_min,
                                 g.data_ptr(), gs[0].data_ptr(), gs[1].data_ptr(), g3)
        return None, None, None, gs[0], gs[1], (gs[2] if len(es) > 2 else None)


class _Cosine(Function):
    """nn.CosineSimilarity(dim=0, eps=1e-8) (model.py:68) or, with ``raw``, the plain dot of
    the BilinearDiag chain score (decoders.py:232)."""

    @staticmethod
    def forward(fctx, x, y, ctx, raw):
        x, y = _c(x), _c(y)
        out = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
        if raw:
            ctx.dot_device(x.size(0), x.size(1), x.data_ptr(), y.data_ptr(), out.data_ptr())
        else:
            ctx.cosine_device(x.size(0), x.size(1), x.data_ptr(), y.data_ptr(), out.data_ptr())
        fctx.save_for_backward(x, y)
        fctx.gqe = (ctx, raw)
        return out

    @staticmethod
    def backward(fctx, g):
        x, y = fctx.saved_tensors
        ctx, raw = fctx.gqe
        g = _c(g)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        ctx.cosine_bwd_device(x.size(0), x.size(1), x.data_ptr(), y.data_ptr(), g.data_ptr(), raw, gx.data_ptr(),
                              gy.data_ptr())
        return gx, gy, None, None


class DifferentiablePath(object):
    """``QueryEncoderDecoder.forward`` (model.py:70-109) as a chain of the Functions above."""

    def __init__(self, model):
        self.m = model
        self.ctx = model.context()
        self.kind = model.path_dec.kind

    # -- operators -------------------------------------------------------------------
    def encode(self, nodes, mode):
        enc = self.m.enc
        rows = torch.from_numpy(np.ascontiguousarray(enc.rows(nodes, mode))).to(self.m.device)
        hook = getattr(self.m, "row_hook", None)
        if hook is not None:
            hook(mode, rows)          # SparseRowAdam: bring these rows up to date before they are read
        return _Encode.apply(enc.table(mode), self.ctx, enc.mode_ids[mode], rows, bool(getattr(self.m, "sparse_table_grads", False)))

    def project(self, embeds, rel):
        dec = self.m.path_dec
        rid = dec.rel_ids[rel]                       # KeyError like the reference's dict lookup
        if self.kind == "bilinear":
            return _Matmul.apply(dec.mats[rel], embeds, self.ctx, False)          # decoders.py:150
        return _VecOp.apply(dec.vecs[rel], embeds, self.ctx, rid, self.kind == "bilinear-diag")

    def path_score(self, embeds1, embeds2, rels):
        dec = self.m.path_dec
        act = embeds1
        for r in rels:
            if self.kind == "bilinear":
                act = _Matmul.apply(dec.mats[r], act, self.ctx, True)             # act.mm(M), decoders.py:145
            else:
                act = _VecOp.apply(dec.vecs[r], act, self.ctx, dec.rel_ids[r], self.kind == "bilinear-diag")
        return _Cosine.apply(act, embeds2, self.ctx, self.kind == "bilinear-diag")

    def intersect(self, e1, e2, mode, e3=None):
        idec = self.m.inter_dec
        use_min = idec.agg == "min"
        if idec.kind.endswith("-simple"):
            return _Aggregate.apply(self.ctx, False, use_min, e1, e2, e3)         # decoders.py:311-319
        pre, post = idec.pre_mats[mode], idec.post_mats[mode]
        hidden = [_Matmul.apply(pre, e, self.ctx, False) for e in (e1, e2, e3) if e is not None]
        comb = _Aggregate.apply(self.ctx, True, use_min, hidden[0], hidden[1], hidden[2] if len(hidden) > 2 else None)
        return _Matmul.apply(post, comb, self.ctx, False)                          # decoders.py:299

    # -- model.py:70-109; ``anchor_nodes[k]`` holds the node ids of anchor slot k, ``targets`` is a
    #    list of node-id sequences, each scored against the ONE query side built here ---
    def scores(self, formula, anchor_nodes, targets):
        qt = formula.query_type
        anchors = lambda k: anchor_nodes[k]
        if qt in CHAIN_TYPES:
            a = self.encode(anchors(0), formula.anchor_modes[0])
            return [self.path_score(self.encode(t, formula.target_mode), a, formula.rels) for t in targets]
        if qt in FLAT_INTER_TYPES or qt == "3-inter_chain":
            e1 = self.project(self.encode(anchors(0), formula.anchor_modes[0]), reverse_relation(formula.rels[0]))
            e2 = self.encode(anchors(1), formula.anchor_modes[1])
            if qt == "3-inter_chain":
                for r in formula.rels[1][::-1]:
                    e2 = self.project(e2, reverse_relation(r))
            else:
                e2 = self.project(e2, reverse_relation(formula.rels[1]))
            e3 = None
            if qt == "3-inter":
                e3 = self.project(self.encode(anchors(2), formula.anchor_modes[2]), reverse_relation(formula.rels[2]))
            q = self.intersect(e1, e2, formula.target_mode, e3)
        elif qt == "3-chain_inter":
            e1 = self.project(self.encode(anchors(0), formula.anchor_modes[0]), reverse_relation(formula.rels[1][0]))
            e2 = self.project(self.encode(anchors(1), formula.anchor_modes[1]), reverse_relation(formula.rels[1][1]))
            q = self.intersect(e1, e2, formula.rels[0][-1])
            q = self.project(q, reverse_relation(formula.rels[0]))
        else:
            return None
        return [_Cosine.apply(self.encode(t, formula.target_mode), q, self.ctx, False) for t in targets]


def anchors_of(formula, queries):
    """Per-slot anchor node ids of a list of Query objects (model.py:75,80,83,91)."""
    n = len(queries)
    return [np.fromiter((q.anchor_nodes[k] for q in queries), dtype=np.int64, count=n)
            for k in range(len(formula.anchor_modes))]


def forward(model, formula, queries, source_nodes):
    out = DifferentiablePath(model).scores(formula, anchors_of(formula, queries), [source_nodes])
    return None if out is None else out[0]


def margin_loss(model, formula, anchor_nodes, pos_nodes, neg_nodes, margin=1):
    """model.py:122-126 with the query side built once for the positive and the negative pass."""
    pos, neg = DifferentiablePath(model).scores(formula, anchor_nodes, [pos_nodes, neg_nodes])
    loss = margin - (pos - neg)
    loss = torch.clamp(loss, min=0)
    return loss.mean()
