"""Host-side query data model: the index producers of the hot path.

Mirrors the three symbols of reference ``netquery/graph.py`` that decide which
node / relation index feeds which operand of the scorer:

* ``reverse_relation``  -- graph.py:4-5
* ``Formula``           -- graph.py:11-36 (structure, relations, modes)
* ``Query``             -- graph.py:38-66 (target, anchors, negative lists)

plus ``QueryBatch``, a flat-array form of "one formula, many queries" that the
CUDA path consumes without touching per-query Python objects.
"""
import random

import numpy as np

CHAIN_TYPES = ("1-chain", "2-chain", "3-chain")
FLAT_INTER_TYPES = ("2-inter", "3-inter")
NESTED_TYPES = ("3-inter_chain", "3-chain_inter")
QUERY_TYPES = CHAIN_TYPES + FLAT_INTER_TYPES + NESTED_TYPES


def reverse_relation(relation):
    """Swap the two modes of a relation triple, keep its name."""
    m_from, name, m_to = relation[0], relation[1], relation[-1]
    return (m_to, name, m_from)


def _anchor_modes(query_type, rels):
    if query_type in CHAIN_TYPES:
        return (rels[-1][-1],)
    if query_type in FLAT_INTER_TYPES:
        return tuple(rel[-1] for rel in rels)
    if query_type == "3-inter_chain":
        return (rels[0][-1], rels[1][-1][-1])
    if query_type == "3-chain_inter":
        return (rels[1][0][-1], rels[1][1][-1])
    return ()


class Formula(object):
    """A query structure with its typed relations; hashable batching key."""

    __slots__ = ("query_type", "rels", "target_mode", "anchor_modes")

    def __init__(self, query_type, rels):
        self.query_type = query_type
        self.rels = rels
        self.target_mode = rels[0][0]
        self.anchor_modes = _anchor_modes(query_type, rels)

    def __hash__(self):
        return hash((self.query_type, self.rels))

    def __eq__(self, other):
        return (self.query_type, self.rels) == (other.query_type, other.rels)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __str__(self):
        return self.query_type + ": " + str(self.rels)

    __repr__ = __str__


def _cap(samples, limit, strict):
    """Keep a sample list as is when short enough, else draw ``limit`` of them.

    ``strict`` reproduces the reference's asymmetry: negatives are kept when
    ``len < limit`` (graph.py:60), hard negatives when ``len <= limit``
    (graph.py:64).
    """
    if samples is None:
        return None
    n = len(samples)
    if (n < limit) if strict else (n <= limit):
        return list(samples)
    return random.sample(list(samples), limit)


def parse_query_graph(query_graph):
    """The index content of one stored query graph (graph.py:40-54):
    -> (query_type, rels, target_node, anchor_nodes).  An edge is
    ``(node_u, (mode_u, rel, mode_v), node_v)``; nested structures hold their
    second branch as a pair of edges."""
    query_type, edges = query_graph[0], query_graph[1:]
    if query_type in CHAIN_TYPES:
        rels = tuple(edge[1] for edge in edges)
        anchors = (edges[-1][-1],)
    elif query_type in FLAT_INTER_TYPES:
        rels = tuple(edge[1] for edge in edges)
        anchors = tuple(edge[-1] for edge in edges)
    elif query_type in NESTED_TYPES:
        first, (second_a, second_b) = edges[0], edges[1]
        rels = (first[1], (second_a[1], second_b[1]))
        anchors = (first[-1], second_b[-1]) if query_type == "3-inter_chain" else (second_a[-1], second_b[-1])
    else:
        raise ValueError("unknown query type %r" % (query_type,))
    return query_type, rels, edges[0][0], anchors


class Query(object):
    """One sampled query graph with its negative samples."""

    def __init__(self, query_graph, neg_samples, hard_neg_samples, neg_sample_max=100, keep_graph=False):
        query_type, rels, self.target_node, self.anchor_nodes = parse_query_graph(query_graph)
        self.formula = Formula(query_type, rels)
        self.query_graph = query_graph if keep_graph else None
        self.neg_samples = _cap(neg_samples, neg_sample_max, strict=True)
        self.hard_neg_samples = _cap(hard_neg_samples, neg_sample_max, strict=False)

    def __hash__(self):
        return hash((self.formula, self.target_node, self.anchor_nodes))

    def __eq__(self, other):
        return (self.formula, self.target_node, self.anchor_nodes) == \
               (other.formula, other.target_node, other.anchor_nodes)

    def serialize(self):
        if self.query_graph is None:
            raise Exception("Cannot serialize query loaded with query graph!")
        return (self.query_graph, self.neg_samples, self.hard_neg_samples)

    @staticmethod
    def deserialize(serial_info, keep_graph=False):
        """graph.py:98-100: the cap is the stored list's own length."""
        graph, negs, hard = serial_info
        return Query(graph, negs, hard, None if negs is None else len(negs), keep_graph=keep_graph)


class QueryBatch(object):
    """Flat node-id arrays for many queries of ONE formula.

    anchors  int64 [n_anchors, n_queries]   node ids
    targets  int64 [n_pairs]                node ids, grouped by query
    offsets  int64 [n_queries + 1] or None  (None: regular layout,
             n_pairs = n_queries * targets_per_query)
    """

    def __init__(self, formula, anchors, targets, offsets=None):
        self.formula = formula
        # node ids are kept as given when they already are int32 (the flat QueryStore's native
        # type and what the node-id entry points of the C ABI take), else widened to int64
        keep32 = getattr(anchors, "dtype", None) == np.int32 and getattr(targets, "dtype", None) == np.int32
        dt = np.int32 if keep32 else np.int64
        self.anchors = np.ascontiguousarray(anchors, dtype=dt)
        self.targets = np.ascontiguousarray(targets, dtype=dt).reshape(-1)
        self.offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
        if self.anchors.ndim != 2 or self.anchors.shape[0] != len(formula.anchor_modes):
            raise ValueError("anchors must be [n_anchors, n_queries]")
        nq = self.anchors.shape[1]
        if self.offsets is None:
            if nq and self.targets.size % nq:
                raise ValueError("regular layout needs len(targets) to be a multiple of n_queries")
        elif self.offsets.shape != (nq + 1,) or (nq + 1 and (self.offsets[0] != 0 or self.offsets[-1] != self.targets.size)):
            raise ValueError("offsets must be [n_queries+1], start at 0 and end at len(targets)")

    def int32_ids(self):
        """(anchors int32 [A,Q], targets int32 [P]) -- the form the node-id entry points of the C
        ABI take -- or None when an id does not fit 32 bits.  Cached."""
        ids = self.__dict__.get("_ids32")
        if ids is None and self.anchors.dtype == np.int32:
            ids = self.__dict__["_ids32"] = (self.anchors, self.targets)
        if ids is None:
            lim = 2 ** 31
            ok = True
            for arr in (self.anchors, self.targets):
                if arr.size and (arr.min() < -lim or arr.max() >= lim):
                    ok = False
            ids = (self.anchors.astype(np.int32), self.targets.astype(np.int32)) if ok else False
            self.__dict__["_ids32"] = ids
        return ids or None

    @property
    def n_queries(self):
        return self.anchors.shape[1]

    @property
    def n_pairs(self):
        return self.targets.size

    @staticmethod
    def from_queries(formula, queries, source_nodes):
        """Lower the reference's ``forward(formula, queries, source_nodes)``
        arguments (model.py:70): pair i scores ``source_nodes[i]`` against
        ``queries[i]``.  Pairs that share a Query object (the eval callers repeat
        each query once per negative, utils.py:58-60,86-88) are grouped so the
        query side is evaluated once.  Returns (batch, order) with
        ``scores_in_call_order[order] = scores_in_batch_order``."""
        n = len(queries)
        if n != len(source_nodes):
            raise ValueError("queries and source_nodes differ in length")
        n_anchor = len(formula.anchor_modes)
        targets = np.fromiter(source_nodes, dtype=np.int64, count=n)
        ids = np.fromiter((id(q) for q in queries), dtype=np.int64, count=n)
        uniq, first, inverse = np.unique(ids, return_index=True, return_inverse=True)
        if len(uniq) == n:
            anchors = np.empty((n_anchor, n), dtype=np.int64)
            for k in range(n_anchor):
                anchors[k] = np.fromiter((q.anchor_nodes[k] for q in queries), dtype=np.int64, count=n)
            return QueryBatch(formula, anchors, targets), None
        # keep first-appearance order of the distinct queries
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
        group = rank[inverse]
        order = np.argsort(group, kind="stable")
        counts = np.bincount(group, minlength=len(uniq))
        offsets = np.zeros(len(uniq) + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        reps = [queries[i] for i in np.sort(first)]
        anchors = np.empty((n_anchor, len(reps)), dtype=np.int64)
        for k in range(n_anchor):
            anchors[k] = np.fromiter((q.anchor_nodes[k] for q in reps), dtype=np.int64, count=len(reps))
        return QueryBatch(formula, anchors, targets[order], offsets), order
