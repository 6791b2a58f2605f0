"""The reference's operator surface, backed by the sm_100a CUDA library.

NOTE (scope): the standalone operator calls here (``enc.forward``, ``path_dec.forward`` /
``.project``, ``inter_dec.forward``) are INFERENCE calls: they detach their inputs and
return tensors without ``grad_fn``.  Training goes through
``graphqembed_b200.QueryEncoderDecoder.margin_loss``, whose differentiable chain
(``autograd.py``) runs the same kernels with hand-written backward passes.

Classes here keep the constructor signatures, method names, attribute dicts
and ``state_dict`` key names of reference ``netquery/encoders.py`` and
``netquery/decoders.py`` so that they can be handed to the reference's own
factories / training scripts:

* ``DirectEncoder(features, feature_modules)``                 encoders.py:11-45
* ``BilinearMetapathDecoder(relations, dims)``                 decoders.py:123-150
* ``TransEMetapathDecoder(relations, dims)``                   decoders.py:181-208
* ``BilinearDiagMetapathDecoder(relations, dims)``             decoders.py:211-236
* ``SetIntersection(mode_dims, expand_dims, agg_func)``        decoders.py:270-300
* ``SimpleSetIntersection(agg_func)``                          decoders.py:302-319

Every ``forward`` / ``project`` is a CUDA kernel launch through the C ABI
(``include/gqe.h``); tensors are feature-major ``[d, B]`` like the
reference's.  Parameters stay ordinary ``nn.Parameter``s (the kernels read
them in place), so optimisers and ``state_dict`` work unchanged.  There is no
CPU path: calling these with CPU parameters raises.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .lowering import RowLookup


def _items(d):
    return d.iteritems() if hasattr(d, "iteritems") else d.items()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s lives on %s; graphqembed_b200 has no CPU path -- move the module to a CUDA device"
                           % (what, t.device))
    return t


def _uniform_dim(dims, what):
    vals = set(int(v) for v in dims.values())
    if len(vals) != 1:
        raise ValueError("%s: per-mode dimensions %s differ; the CUDA path needs one embedding dimension" % (what, dims))
    return vals.pop()


def _fm(t, d):
    """Validate a feature-major [d, n] fp32 CUDA tensor, make it contiguous."""
    if t.dim() != 2 or t.size(0) != d:
        raise ValueError("expected a [%d, n] feature-major tensor, got %s" % (d, tuple(t.shape)))
    _require_cuda(t, "embedding tensor")
    return t.detach().float().contiguous()


class _CudaOperator(nn.Module):
    """Lazily owns a gqe context bound to this module's own parameters."""

    def _first_param(self):
        return next(self.parameters())

    def _ctx(self):
        p = _require_cuda(self._first_param(), type(self).__name__ + " parameters")
        dev = p.device.index if p.device.index is not None else torch.cuda.current_device()
        state = self.__dict__.get("_gqe_state")
        sig = self._pointer_signature()
        if state is None or state[1] != dev:
            state = [_lib.Context(dev), dev, None]
            self.__dict__["_gqe_state"] = state
        if state[2] != sig:
            self._bind(state[0])
            state[2] = sig
        state[0].set_stream(torch.cuda.current_stream(dev).cuda_stream)
        return state[0]

    def _pointer_signature(self):
        return tuple(p.data_ptr() for p in self.parameters())

    def _bind(self, ctx):
        raise NotImplementedError

    def __getstate__(self):
        # the native context (a ctypes handle) is per process: recreated lazily after
        # copy.deepcopy / pickle / torch.save(module)
        state = dict(self.__dict__)
        state.pop("_gqe_state", None)
        return state


class DirectEncoder(_CudaOperator):
    """Embedding lookup + L2 normalisation (reference encoders.py:11-45).

    ``features`` should be a ``RowLookup`` (node id -> table row); the lookup
    and the normalisation then run in one gather kernel.  ``feature_modules``
    maps mode -> ``nn.Embedding`` and is registered as ``feat-<mode>`` exactly
    like the reference, so checkpoints are interchangeable.
    """

    def __init__(self, features, feature_modules):
        super(DirectEncoder, self).__init__()
        self.modes = []
        for name, module in _items(feature_modules):
            self.add_module("feat-" + name, module)
            self.modes.append(name)
        self.mode_ids = {m: i for i, m in enumerate(self.modes)}
        self.feature_modules = dict(_items(feature_modules))
        if not isinstance(features, RowLookup):
            raise TypeError(
                "DirectEncoder needs a graphqembed_b200.RowLookup as `features` (node id -> table row); an opaque "
                "embedding closure cannot be fused into the CUDA gather. Build one from load_graph's node_maps.")
        self.features = features
        self.dim = _uniform_dim({m: mod.weight.size(1) for m, mod in self.feature_modules.items()}, "DirectEncoder")

    def table(self, mode):
        return self.feature_modules[mode].weight

    def _bind(self, ctx):
        ws = [_require_cuda(self.table(m), "embedding table").detach() for m in self.modes]
        for w in ws:
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise ValueError("embedding tables must be contiguous fp32")
        ctx.bind_tables([w.data_ptr() for w in ws], [w.size(0) for w in ws], self.dim)

    def rows(self, nodes, mode):
        return self.features.rows(nodes, mode)

    def forward(self, nodes, mode, offset=None, **kwargs):
        if offset is not None:
            raise NotImplementedError("EmbeddingBag-style offsets (reference encoders.py:44-45, Reddit data) are "
                                      "outside the accelerated path")
        ctx = self._ctx()
        rows = torch.from_numpy(self.rows(nodes, mode)).to(self.table(mode).device, non_blocking=True)
        out = torch.empty((self.dim, rows.numel()), dtype=torch.float32, device=rows.device)
        ctx.encode_device(self.mode_ids[mode], rows.numel(), rows.data_ptr(), out.data_ptr())
        return out


class _MetapathDecoder(_CudaOperator):
    kind = None
    matrix = False

    def __init__(self, relations, dims):
        super(_MetapathDecoder, self).__init__()
        self.relations = relations
        self.dim = _uniform_dim(dims, type(self).__name__)
        self.rel_keys = []
        store = {}
        for r1 in relations:
            for r2 in relations[r1]:
                rel = (r1, r2[1], r2[0])
                if self.matrix:
                    p = nn.Parameter(torch.empty(dims[rel[0]], dims[rel[2]]))
                    nn.init.xavier_uniform_(p)
                else:
                    bound = 6.0 / math.sqrt(dims[rel[0]])
                    p = nn.Parameter(torch.empty(dims[rel[0]]))
                    nn.init.uniform_(p, a=-bound, b=bound)
                store[rel] = p
                self.register_parameter("_".join(rel), p)
                self.rel_keys.append(rel)
        self.rel_ids = {rel: i for i, rel in enumerate(self.rel_keys)}
        if self.matrix:
            self.mats = store
        else:
            self.vecs = store

    def _store(self):
        return self.mats if self.matrix else self.vecs

    def _apply(self, fn, *args, **kwargs):
        # nn.Module._apply may replace Parameter objects' data; keep the dict views pointing at the registered ones
        out = super(_MetapathDecoder, self)._apply(fn, *args, **kwargs)
        store = self._store()
        for rel in self.rel_keys:
            store[rel] = self._parameters["_".join(rel)]
        return out

    def _bind(self, ctx):
        ps = [self._store()[rel].detach() for rel in self.rel_keys]
        for p in ps:
            _require_cuda(p, "relation parameter")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("relation parameters must be contiguous fp32")
        ctx.bind_relations(_lib.DECODER_ID[self.kind], [p.data_ptr() for p in ps], self.dim)

    def forward(self, embeds1, embeds2, rels):
        """Score of a metapath between two embedding batches -> [B]."""
        ctx = self._ctx()
        ids = [self.rel_ids[r] for r in rels]          # KeyError like self.mats[i_rel]
        _require_cuda(embeds1, "embeds1")
        if embeds1.dim() != 2 or embeds1.size(0) != self.dim:
            raise ValueError("expected a [%d, n] feature-major tensor, got %s" % (self.dim, tuple(embeds1.shape)))
        mutate = self.kind == "transe" and embeds1.is_contiguous() and embeds1.dtype == torch.float32
        e1 = embeds1.detach() if mutate else _fm(embeds1, self.dim)
        e2 = _fm(embeds2, self.dim)
        if e1.shape != e2.shape:
            raise ValueError("embeds1 %s and embeds2 %s differ in shape" % (tuple(e1.shape), tuple(e2.shape)))
        n = e1.size(1)
        out = torch.empty(n, dtype=torch.float32, device=e1.device)
        # reference TransE translates embeds1 in place (decoders.py:203); reproduced when it can alias
        ctx.path_score_device(ids, n, e1.data_ptr(), e2.data_ptr(), mutate, out.data_ptr())
        return out

    def project(self, embeds, rel):
        ctx = self._ctx()
        rid = self.rel_ids[rel]
        e = _fm(embeds, self.dim)
        out = torch.empty_like(e)
        ctx.project_device(rid, e.size(1), e.data_ptr(), out.data_ptr())
        return out


class BilinearMetapathDecoder(_MetapathDecoder):
    """One d x d matrix per relation; a metapath is a product of matrices."""
    kind = "bilinear"
    matrix = True


class TransEMetapathDecoder(_MetapathDecoder):
    """One translation vector per relation; a metapath is a sum of vectors."""
    kind = "transe"
    matrix = False


class BilinearDiagMetapathDecoder(_MetapathDecoder):
    """DistMult: one diagonal per relation; the chain score is a raw dot product."""
    kind = "bilinear-diag"
    matrix = False


def _agg_name(agg_func):
    if agg_func in ("mean", "min"):
        return agg_func
    if agg_func is torch.mean:
        return "mean"
    if agg_func is torch.min:
        return "min"
    raise ValueError("agg_func must be torch.mean, torch.min, 'mean' or 'min' (got %r)" % (agg_func,))


class SetIntersection(_CudaOperator):
    """DeepSets intersection: post[mode] . agg_k relu(pre[mode] . e_k)."""

    def __init__(self, mode_dims, expand_dims, agg_func=torch.min):
        super(SetIntersection, self).__init__()
        self.agg_func = agg_func
        self.agg = _agg_name(agg_func)
        self.dim = _uniform_dim(mode_dims, "SetIntersection")
        self.expand_dim = _uniform_dim(expand_dims, "SetIntersection expand")
        self.modes = list(mode_dims)
        self.mode_ids = {m: i for i, m in enumerate(self.modes)}
        self.pre_mats = {}
        self.post_mats = {}
        for mode in self.modes:
            pre = nn.Parameter(torch.empty(expand_dims[mode], mode_dims[mode]))
            nn.init.xavier_uniform_(pre)
            self.register_parameter(mode + "_premat", pre)
            self.pre_mats[mode] = pre
            post = nn.Parameter(torch.empty(mode_dims[mode], expand_dims[mode]))
            nn.init.xavier_uniform_(post)
            self.register_parameter(mode + "_postmat", post)
            self.post_mats[mode] = post

    @property
    def kind(self):
        return self.agg

    def _apply(self, fn, *args, **kwargs):
        out = super(SetIntersection, self)._apply(fn, *args, **kwargs)
        for mode in self.modes:
            self.pre_mats[mode] = self._parameters[mode + "_premat"]
            self.post_mats[mode] = self._parameters[mode + "_postmat"]
        return out

    def _bind(self, ctx):
        pre = [_require_cuda(self.pre_mats[m], "pre matrix").detach() for m in self.modes]
        post = [_require_cuda(self.post_mats[m], "post matrix").detach() for m in self.modes]
        ctx.bind_intersection(_lib.INTER_ID[self.kind], [p.data_ptr() for p in pre], [p.data_ptr() for p in post],
                              self.dim, self.expand_dim)

    def forward(self, embeds1, embeds2, mode, embeds3=[]):
        ctx = self._ctx()
        e1, e2 = _fm(embeds1, self.dim), _fm(embeds2, self.dim)
        e3 = _fm(embeds3, self.dim) if len(embeds3) > 0 else None
        out = torch.empty_like(e1)
        ctx.intersect_device(self.mode_ids[mode], e1.size(1), e1.data_ptr(), e2.data_ptr(),
                             None if e3 is None else e3.data_ptr(), out.data_ptr())
        return out


class SimpleSetIntersection(nn.Module):
    """Parameter-free elementwise mean / min of the operands."""

    def __init__(self, agg_func=torch.min):
        super(SimpleSetIntersection, self).__init__()
        self.agg_func = agg_func
        self.agg = _agg_name(agg_func)

    @property
    def kind(self):
        return self.agg + "-simple"

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_gqe_state", None)
        return state

    def forward(self, embeds1, embeds2, mode, embeds3=[]):
        d = embeds1.size(0)
        e1, e2 = _fm(embeds1, d), _fm(embeds2, d)
        e3 = _fm(embeds3, d) if len(embeds3) > 0 else None
        dev = e1.device.index if e1.device.index is not None else torch.cuda.current_device()
        state = self.__dict__.get("_gqe_state")
        if state is None or state[1] != (dev, d):
            ctx = _lib.Context(dev)
            ctx.bind_intersection(_lib.INTER_ID[self.kind], None, None, d)
            state = (ctx, (dev, d))
            self.__dict__["_gqe_state"] = state
        ctx = state[0]
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        out = torch.empty_like(e1)
        ctx.intersect_device(0, e1.size(1), e1.data_ptr(), e2.data_ptr(), None if e3 is None else e3.data_ptr(),
                             out.data_ptr())
        return out


def cosine_similarity_dim0(x, y):
    """nn.CosineSimilarity(dim=0, eps=1e-8) on feature-major CUDA tensors (model.py:68)."""
    d = x.size(0)
    x, y = _fm(x, d), _fm(y, d)
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    ctx = _SHARED_CTX.get(dev)
    if ctx is None:
        ctx = _SHARED_CTX[dev] = _lib.Context(dev)     # parameter-free calls share one context per device
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    out = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
    ctx.cosine_device(d, x.size(1), x.data_ptr(), y.data_ptr(), out.data_ptr())
    return out


_SHARED_CTX = {}


# ---- factories with the reference's names (netquery/utils.py:93-150) -------------
def get_encoder(depth, graph, out_dims, feature_modules, cuda=True):
    """utils.py:93-126: depth 0 = DirectEncoder (the fused path), 1..3 = stacked GraphSAGE-style
    Encoders (``sage.py``; un-fused operator chain)."""
    from .sage import get_encoder as _get
    return _get(depth, graph, out_dims, feature_modules, cuda)


def get_metapath_decoder(graph, out_dims, decoder):
    if decoder == "bilinear":
        return BilinearMetapathDecoder(graph.relations, out_dims)
    if decoder == "transe":
        return TransEMetapathDecoder(graph.relations, out_dims)
    if decoder == "bilinear-diag":
        return BilinearDiagMetapathDecoder(graph.relations, out_dims)
    raise Exception("Metapath decoder not recognized.")


def get_intersection_decoder(graph, out_dims, decoder):
    if decoder == "mean":
        return SetIntersection(out_dims, out_dims, agg_func=torch.mean)
    if decoder == "mean-simple":
        return SimpleSetIntersection(agg_func=torch.mean)
    if decoder == "min":
        return SetIntersection(out_dims, out_dims, agg_func=torch.min)
    if decoder == "min-simple":
        return SimpleSetIntersection(agg_func=torch.min)
    raise Exception("Intersection decoder not recognized.")
