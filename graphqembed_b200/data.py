"""On-disk formats and batch sampling on either side of the scoring path.

Two layers:

* ``store.QueryStore`` (re-exported here) -- the native form: a query file parsed once
  into flat per-formula int32 arrays; batches are views, negatives are drawn
  vectorised, the scorer takes the slices directly.  This is what a training /
  evaluation loop built on this package should hold.
* the reference's loader NAMES (netquery/data_utils.py:6-35, netquery/bio/data_utils.py:
  11-23, netquery/train_helpers.py:95-107), for scripts written against netquery:
  they return the same nestings of ``Query`` objects the reference returns, built by
  one generic grouping pass.

The files are the reference's own pickles (written by Python 2, hence
``encoding="latin1"``): ``graph_data.pkl = (rels, adj_lists, node_maps)`` and query
files = lists of ``(query_graph, neg_samples, hard_neg_samples)`` (graph.py:93-100).
Nothing here computes scores.
"""
import pickle

import numpy as np

from .lowering import RowLookup
from .query import Query
from .store import FormulaBlock, QueryStore, StoreSlice, batch_window  # noqa: F401  (public)


def _records(path):
    with open(path, "rb") as fh:
        return pickle.load(fh, encoding="latin1")


def _nested(pairs, depth):
    """[(key tuple of length ``depth``, value)] -> dicts nested ``depth`` deep with lists at
    the leaves, keys in first-appearance order."""
    root = {}
    for keys, value in pairs:
        node = root
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node.setdefault(keys[-1], []).append(value)
    return root


class _Missing(dict):
    """The reference hands out ``defaultdict``s: an absent query type reads as empty."""

    def __init__(self, src, leaf):
        dict.__init__(self, src)
        self._leaf = leaf

    def __missing__(self, key):
        return self._leaf()


class GraphData(object):
    """The slice of reference ``Graph`` (netquery/graph.py:104-121) the scoring path reads:
    ``relations`` (decoder registration order), ``full_lists`` (1-chain negatives,
    model.py:118), ``features`` (the row lookup) and ``feature_dims``; ``adj_lists`` is kept
    for callers that sample."""

    def __init__(self, features, feature_dims, relations, adj_lists):
        self.features, self.feature_dims = features, feature_dims
        self.relations, self.adj_lists = relations, adj_lists
        # every node with at least one outgoing edge, per source mode (graph.py:116-120)
        self.full_sets = {}
        for (src_mode, _, _), by_node in adj_lists.items():
            self.full_sets.setdefault(src_mode, set()).update(by_node)
        self.full_lists = {mode: list(nodes) for mode, nodes in self.full_sets.items()}
        self._full_arrays = {}

    def full_array(self, mode):
        """``full_lists[mode]`` as an int32 array (vectorised 1-chain negatives)."""
        arr = self._full_arrays.get(mode)
        if arr is None:
            arr = self._full_arrays[mode] = np.asarray(self.full_lists[mode], dtype=np.int32)
        return arr


def load_graph(data_dir, embed_dim, graph_file="graph_data.pkl"):
    """bio/data_utils.py:11-23 -> (graph, feature_modules, node_maps).

    ``node_maps[mode]`` maps node id -> position and carries the reference's extra
    ``-1 -> -1`` entry (:14-15), so a table has N + 2 rows (:16-17), initialised N(0, 1/d)
    (:18-19).  ``graph.features`` is a ``RowLookup`` (row = ``node_maps[mode][n] + 1``, :20-21)
    instead of an embedding closure: the CUDA path gathers the rows itself."""
    import torch
    relations, adj_lists, ids_by_mode = _records(data_dir + "/" + graph_file)
    node_maps = {}
    for mode, ids in ids_by_mode.items():
        node_maps[mode] = dict(zip(ids, range(len(ids))))
        node_maps[mode][-1] = -1
    feature_modules = {}
    for mode in relations:
        table = torch.nn.Embedding(len(node_maps[mode]) + 1, embed_dim)
        table.weight.data.normal_(0, 1. / embed_dim)
        feature_modules[mode] = table
    graph = GraphData(RowLookup(node_maps), {mode: embed_dim for mode in relations}, relations, adj_lists)
    return graph, feature_modules, node_maps


# ---- the reference's loader names ---------------------------------------------------
def load_queries(data_file, keep_graph=False):
    """data_utils.py:6-8 -> [Query]"""
    return [Query.deserialize(rec, keep_graph=keep_graph) for rec in _records(data_file)]


def load_queries_by_formula(data_file):
    """data_utils.py:10-16 -> {query_type: {Formula: [Query]}}"""
    qs = load_queries(data_file)
    return _Missing(_nested((((q.formula.query_type, q.formula), q) for q in qs), 2), dict)


def load_queries_by_type(data_file, keep_graph=True):
    """data_utils.py:18-24 -> {query_type: [Query]}"""
    qs = load_queries(data_file, keep_graph=keep_graph)
    return _Missing(_nested((((q.formula.query_type,), q) for q in qs), 1), list)


def load_test_queries_by_formula(data_file):
    """data_utils.py:27-35 -> {"full_neg" | "one_neg": {query_type: {Formula: [Query]}}}, split
    by whether more than one negative was stored with the query."""
    keyed = []
    for rec in _records(data_file):
        q = Query.deserialize(rec)
        keyed.append((("full_neg" if len(rec[1]) > 1 else "one_neg", q.formula.query_type, q.formula), q))
    out = _nested(keyed, 3)
    return {split: _Missing(out.get(split, {}), dict) for split in ("full_neg", "one_neg")}


def pick_batch(train_queries, iter_count, batch_size):
    """The batch ``run_batch`` scores (train_helpers.py:96-105): ONE formula drawn
    ~ multinomial(#queries per formula) from numpy's global RNG, then a contiguous, wrapping
    window of its queries.  ``train_queries``: {Formula: list-like of queries} (lists of
    ``Query``, ``FormulaBlock``s or ``DeviceBlock``s).  -> (formula, queries)"""
    formulas = list(train_queries)
    sizes = np.array([len(train_queries[f]) for f in formulas], dtype=np.float64)
    formula = formulas[int(np.argmax(np.random.multinomial(1, sizes / sizes.sum())))]
    pool = train_queries[formula]
    start, stop = batch_window(iter_count, batch_size, len(pool))
    return formula, (pool.window(start, stop) if hasattr(pool, "window") else pool[start:stop])   # store blocks (host or device)


def run_batch(train_queries, enc_dec, iter_count, batch_size, hard_negatives=False):
    """train_helpers.py:95-107"""
    formula, queries = pick_batch(train_queries, iter_count, batch_size)
    return enc_dec.margin_loss(formula, queries, hard_negatives=hard_negatives)
