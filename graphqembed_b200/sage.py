"""GraphSAGE-style node encoder: the reference's ``--depth > 0`` path.

Drop-ins for ``netquery/aggregators.py:17-68`` (``MeanAggregator``) and
``netquery/encoders.py:47-129`` (``Encoder``) as ``netquery/utils.py:93-126`` stacks them:

    self_feat  = features(nodes, mode)                                   raw rows, NOT normalised
    to_feats_r = mean over sampled neighbours under relation r of features(neighbour)
    out        = relu( compress[mode] . concat(to_feats_r1, ..., to_feats_rn, self_feat) )   [d_out, B]

Same constructor arguments, attribute / state-dict names (``feat-<mode>``, ``<mode>_compress``)
and the same neighbour sampling -- ``random.sample`` per node from the global ``random`` stream, so
a seeded run draws the reference's neighbours.  What changes is where the arithmetic runs: the
sampled lists become a CSR, the masked mean is one gather-reduce kernel over the embedding table
(or over a lower encoder's output), the compression one fp32 GEMM + ReLU kernel
(``csrc/gqe_sage.cu`` behind ``gqe_segment_mean_device`` / ``gqe_linear_device``).  The dense
[batch, unique-neighbours] mask of the reference is never built.

A model whose encoder is one of these runs the un-fused operator chain (``QueryEncoderDecoder``
detects it); the fused kernels need the ``DirectEncoder``.  Inference calls: no ``grad_fn``.
"""
import math
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .lowering import RowLookup
from .operators import DirectEncoder, _CudaOperator, _items, _require_cuda


def sample_neighbours(to_neighs, keep_prob=0.5, max_keep=10):
    """aggregators.py:50-53: per node ``min(ceil(len * keep_prob), max_keep)`` distinct neighbours
    drawn with ``random.sample`` (global stream, node order).
    -> (ptr int64 [n+1], flat node list, the per-node sets)"""
    ptr = np.zeros(len(to_neighs) + 1, dtype=np.int64)
    flat, sets = [], []
    for i, neigh in enumerate(to_neighs):
        neigh = list(neigh)
        picked = set(random.sample(neigh, min(int(math.ceil(len(neigh) * keep_prob)), max_keep)))
        sets.append(picked)
        flat.extend(picked)
        ptr[i + 1] = len(flat)
    return ptr, flat, sets


class MeanAggregator(_CudaOperator):
    """Mean of sampled neighbours' features (aggregators.py:17-68).

    ``features``: a ``RowLookup`` together with ``feature_modules`` (mode -> nn.Embedding) -- the
    rows are then gathered straight from the table -- or any callable ``(nodes, mode) -> [n, d]``
    CUDA tensor (a lower ``Encoder``, as utils.py:108,117 wires it)."""

    def __init__(self, features, cuda=False, feature_modules=None):
        super(MeanAggregator, self).__init__()
        self.features = features
        self.cuda = cuda
        self.__dict__["_tables"] = dict(_items(feature_modules)) if feature_modules else None   # not re-registered

    def _ctx_for(self, device):
        dev = device.index if device.index is not None else torch.cuda.current_device()
        state = self.__dict__.get("_gqe_state")
        if state is None or state[1] != dev:
            state = [_lib.Context(dev), dev, None]
            self.__dict__["_gqe_state"] = state
        state[0].set_stream(torch.cuda.current_stream(dev).cuda_stream)
        return state[0]

    def forward(self, to_neighs, rel, keep_prob=0.5, max_keep=10):
        """-> [len(to_neighs), d] (the reference's ``to_feats``)."""
        ptr, flat, sets = sample_neighbours(to_neighs, keep_prob, max_keep)
        mode = rel[-1]
        if isinstance(self.features, RowLookup) and self._tables is not None:
            src = _require_cuda(self._tables[mode].weight, "embedding table").detach()
            cols = self.features.rows(flat, mode)                      # table rows (-1 -> row 0, bio/data_utils.py:15)
        else:
            # the lower encoder samples ITS neighbours from the same random stream: it must see the
            # unique neighbours in the reference's order, list(set.union(*samp_neighs)) (aggregators.py:54)
            uniq = list(set.union(*sets))
            pos = {n: i for i, n in enumerate(uniq)}
            inverse = np.fromiter((pos[n] for n in flat), dtype=np.int32, count=len(flat))
            src = self.features(uniq, mode)                            # [n_unique, d] from the lower encoder
            if src.dim() == 1:
                src = src.unsqueeze(0)
            src = _require_cuda(src, "neighbour features").detach().float().contiguous()
            cols = inverse.astype(np.int32)
        n, d = len(to_neighs), src.size(1)
        dev = src.device
        ptr_t = torch.from_numpy(ptr).to(dev)
        cols_t = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int32)).to(dev)
        out = torch.empty((n, d), dtype=torch.float32, device=dev)
        ctx = self._ctx_for(dev)
        ctx.segment_mean_device(src.data_ptr(), src.size(0), d, n, ptr_t.data_ptr(), cols_t.data_ptr(), out.data_ptr())
        return out


class Encoder(_CudaOperator):
    """encoders.py:47-129.  ``features`` / ``aggregator`` as in the reference; ``layer_norm`` is not
    reachable from the reference's factories (utils.py:93-126 never sets it) and is not accelerated."""

    def __init__(self, features, feature_dims, out_dims, relations, adj_lists, aggregator, base_model=None, cuda=False,
                 layer_norm=False, feature_modules={}):
        super(Encoder, self).__init__()
        if layer_norm:
            raise NotImplementedError("layer_norm=True is outside the accelerated path (never set by netquery/utils.py)")
        self.features, self.feat_dims, self.adj_lists = features, feature_dims, adj_lists
        self.relations, self.aggregator = relations, aggregator
        self.modes = []
        for name, module in _items(feature_modules):
            self.add_module("feat-" + name, module)
            self.modes.append(name)
        self.feature_modules = dict(_items(feature_modules))
        if base_model is not None:
            self.base_model = base_model
        self.out_dims, self.cuda = out_dims, cuda
        self.aggregator.cuda = cuda
        self.compress_dims, self.compress_params = {}, {}
        for source_mode in relations:
            self.compress_dims[source_mode] = self.feat_dims[source_mode]
            for (to_mode, _) in relations[source_mode]:
                self.compress_dims[source_mode] += self.feat_dims[to_mode]
        for mode in _items(self.feat_dims):
            mode = mode[0]
            w = nn.Parameter(torch.empty(out_dims[mode], self.compress_dims[mode]))
            nn.init.xavier_uniform_(w)
            self.register_parameter(mode + "_compress", w)
            self.compress_params[mode] = w
        self.dim = int(next(iter(out_dims.values())))

    def _apply(self, fn, *args, **kwargs):
        out = super(Encoder, self)._apply(fn, *args, **kwargs)
        for mode in self.compress_params:
            self.compress_params[mode] = self._parameters[mode + "_compress"]
        return out

    def _bind(self, ctx):
        pass

    def _self_features(self, nodes, mode):
        if isinstance(self.features, RowLookup) and mode in self.feature_modules:
            table = _require_cuda(self.feature_modules[mode].weight, "embedding table").detach()
            rows = torch.from_numpy(self.features.rows(nodes, mode)).to(table.device)
            out = torch.empty((rows.numel(), table.size(1)), dtype=torch.float32, device=table.device)
            ctx = self._ctx()
            # a raw gather is the aggregation of one-element segments
            ptr = torch.arange(rows.numel() + 1, dtype=torch.int64, device=table.device)
            ctx.segment_mean_device(table.data_ptr(), table.size(0), table.size(1), rows.numel(), ptr.data_ptr(),
                                    rows.data_ptr(), out.data_ptr())
            return out
        feats = self.features(nodes, mode)
        if feats.dim() == 1:
            feats = feats.unsqueeze(0)
        return _require_cuda(feats, "features").detach().float()

    def forward(self, nodes, mode, keep_prob=0.5, max_keep=10):
        """-> [d_out, len(nodes)], ReLU'd, not normalised (encoders.py:103-123)."""
        nodes = [int(n) for n in nodes]
        # own features FIRST, as encoders.py:110 does: when they come from a lower encoder that call
        # draws neighbour samples, and the position in the random stream is part of the contract
        self_feat = self._self_features(nodes, mode).t()
        parts = []
        for to_r in self.relations[mode]:
            rel = (mode, to_r[1], to_r[0])
            adj = self.adj_lists[rel]
            to_neighs = [[-1] if node == -1 else adj[node] for node in nodes]
            to_neighs = [[-1] if len(l) == 0 else l for l in to_neighs]        # null neighbour (encoders.py:112-113)
            parts.append(self.aggregator.forward(to_neighs, rel, keep_prob, max_keep).t())
        parts.append(self_feat)
        combined = torch.cat(parts, dim=0).contiguous()                        # [compress_dim, B]
        w = _require_cuda(self.compress_params[mode], "compress matrix").detach()
        out = torch.empty((w.size(0), combined.size(1)), dtype=torch.float32, device=combined.device)
        self._ctx().linear_device(w.data_ptr(), w.size(0), w.size(1), combined.size(1), combined.data_ptr(), 1,
                                  out.data_ptr())
        return out


def get_encoder(depth, graph, out_dims, feature_modules, cuda=True):
    """netquery/utils.py:93-126: depth 0 = DirectEncoder, depth 1..3 = stacked Encoders, each layer's
    features the transposed output of the layer below."""
    if depth < 0 or depth > 3:
        raise Exception("Depth must be between 0 and 3 (inclusive)")
    if depth == 0:
        return DirectEncoder(graph.features, feature_modules)
    agg1 = MeanAggregator(graph.features, feature_modules=feature_modules)
    enc1 = Encoder(graph.features, graph.feature_dims, out_dims, graph.relations, graph.adj_lists,
                   feature_modules=feature_modules, cuda=cuda, aggregator=agg1)
    enc = enc1
    if depth >= 2:
        lower1 = lambda nodes, mode: enc1(nodes, mode).t()
        enc2 = Encoder(lower1, enc1.out_dims, out_dims, graph.relations, graph.adj_lists, base_model=enc1, cuda=cuda,
                       aggregator=MeanAggregator(lower1))
        enc = enc2
        if depth >= 3:
            # (the reference wires layer 3's own features to enc1 and its aggregator to enc2: utils.py:116-119)
            lower2 = lambda nodes, mode: enc2(nodes, mode).t()
            enc = Encoder(lower1, enc2.out_dims, out_dims, graph.relations, graph.adj_lists, base_model=enc2, cuda=cuda,
                          aggregator=MeanAggregator(lower2))
    return enc
