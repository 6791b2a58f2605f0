"""On-disk formats and batch sampling on either side of the scoring path.

Mirrors, for Python 3, the loaders and the batch sampler the reference's train
script calls around ``QueryEncoderDecoder``:

* ``load_graph``                    netquery/bio/data_utils.py:11-23
* ``load_queries`` / ``load_queries_by_formula`` / ``load_queries_by_type`` /
  ``load_test_queries_by_formula``  netquery/data_utils.py:6-35
* ``run_batch``                     netquery/train_helpers.py:95-107

The files are the reference's own pickles (written by Python 2: loaded with
``encoding="latin1"``): ``graph_data.pkl = (rels, adj_lists, node_maps)`` and
query files = lists of ``(query_graph, neg_samples, hard_neg_samples)``
(netquery/graph.py:93-100).  Nothing here computes scores.
"""
import pickle
from collections import defaultdict

import numpy as np

from .lowering import RowLookup
from .query import Query


def _load_pickle(path):
    with open(path, "rb") as fh:
        return pickle.load(fh, encoding="latin1")


class GraphData(object):
    """The slice of reference ``Graph`` (netquery/graph.py:104-121) the scoring
    path reads: ``relations`` (decoder registration order), ``full_lists``
    (1-chain negatives, model.py:118), ``features`` (the row lookup) and
    ``feature_dims``.  ``adj_lists`` is kept for callers that sample."""

    def __init__(self, features, feature_dims, relations, adj_lists):
        self.features = features
        self.feature_dims = feature_dims
        self.relations = relations
        self.adj_lists = adj_lists
        full_sets = defaultdict(set)
        for rel in adj_lists:                                   # graph.py:116-118
            full_sets[rel[0]] = full_sets[rel[0]].union(set(adj_lists[rel].keys()))
        self.full_sets = full_sets
        self.full_lists = {mode: list(s) for mode, s in full_sets.items()}   # graph.py:119-120


def load_graph(data_dir, embed_dim, graph_file="graph_data.pkl"):
    """bio/data_utils.py:11-23 -> (graph, feature_modules, node_maps).

    One ``nn.Embedding(N_mode + 2, d)`` per mode (``node_maps[m][-1] = -1`` adds the
    extra entry, :14-16), initialised N(0, 1/d) (:17-19).  ``graph.features`` is a
    ``RowLookup`` (row = ``node_maps[mode][n] + 1``, :20-21) instead of an embedding
    closure: the CUDA path gathers the rows itself."""
    import torch
    rels, adj_lists, node_ids = _load_pickle(data_dir + "/" + graph_file)
    node_maps = {m: {n: i for i, n in enumerate(id_list)} for m, id_list in node_ids.items()}
    for m in node_maps:
        node_maps[m][-1] = -1
    feature_dims = {m: embed_dim for m in rels}
    feature_modules = {m: torch.nn.Embedding(len(node_maps[m]) + 1, embed_dim) for m in rels}
    for mode in rels:
        feature_modules[mode].weight.data.normal_(0, 1. / embed_dim)
    graph = GraphData(RowLookup(node_maps), feature_dims, rels, adj_lists)
    return graph, feature_modules, node_maps


def load_queries(data_file, keep_graph=False):
    """data_utils.py:6-8"""
    return [Query.deserialize(info, keep_graph=keep_graph) for info in _load_pickle(data_file)]


def load_queries_by_formula(data_file):
    """data_utils.py:10-16 -> {query_type: {Formula: [Query]}}"""
    queries = defaultdict(lambda: defaultdict(list))
    for raw_query in _load_pickle(data_file):
        query = Query.deserialize(raw_query)
        queries[query.formula.query_type][query.formula].append(query)
    return queries


def load_queries_by_type(data_file, keep_graph=True):
    """data_utils.py:18-24 -> {query_type: [Query]}"""
    queries = defaultdict(list)
    for raw_query in _load_pickle(data_file):
        query = Query.deserialize(raw_query, keep_graph=keep_graph)
        queries[query.formula.query_type].append(query)
    return queries


def load_test_queries_by_formula(data_file):
    """data_utils.py:27-35: split by whether more than one negative was stored."""
    queries = {"full_neg": defaultdict(lambda: defaultdict(list)),
               "one_neg": defaultdict(lambda: defaultdict(list))}
    for raw_query in _load_pickle(data_file):
        neg_type = "full_neg" if len(raw_query[1]) > 1 else "one_neg"
        query = Query.deserialize(raw_query)
        queries[neg_type][query.formula.query_type][query.formula].append(query)
    return queries


def pick_batch(train_queries, iter_count, batch_size):
    """The batch ``run_batch`` scores (train_helpers.py:96-105): ONE formula drawn
    ~ multinomial(#queries per formula) from numpy's global RNG, then a
    contiguous, wrapping slice of its queries.  -> (formula, [Query])"""
    formulas = list(train_queries.keys())
    num_queries = [float(len(train_queries[f])) for f in formulas]
    denom = float(sum(num_queries))
    formula_index = np.argmax(np.random.multinomial(1, np.array(num_queries) / denom))
    formula = formulas[formula_index]
    n = len(train_queries[formula])
    start = (iter_count * batch_size) % n
    end = min(((iter_count + 1) * batch_size) % n, n)
    end = n if end <= start else end
    return formula, train_queries[formula][start:end]


def run_batch(train_queries, enc_dec, iter_count, batch_size, hard_negatives=False):
    """train_helpers.py:95-107"""
    formula, queries = pick_batch(train_queries, iter_count, batch_size)
    return enc_dec.margin_loss(formula, queries, hard_negatives=hard_negatives)
