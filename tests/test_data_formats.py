"""Loaders for the reference's on-disk formats and the run_batch sampler
(graphqembed_b200/data.py) against files written the way the reference writes
them (protocol-2 pickles of plain tuples/dicts/lists, netquery/graph.py:93-96,
netquery/bio/data_utils.py:12) and, when the reference tree is mounted,
against the reference's own code."""
import pickle
import random

import numpy as np
import pytest
import torch

import graphqembed_b200 as gqe
from graphqembed_b200 import data
from graphqembed_b200.synth import SynthKG
from oracle import ref_shim
from oracle.cases import make_case


def _write_dataset(tmp_path, case, n_neg_keep):
    kg = case.kg
    rels = kg.relations
    adj = {}
    rng = np.random.RandomState(0)
    for r in kg.rel_keys:                     # {(m1, rel, m2): {node: [neighbours]}}
        src = kg.node_ids[r[0]][:40]
        adj[r] = {int(n): [int(x) for x in kg.node_ids[r[2]][rng.randint(0, 30, size=3)]] for n in src}
    node_ids = {m: [int(n) for n in kg.node_ids[m]] for m in kg.modes}
    with open(tmp_path / "graph_data.pkl", "wb") as fh:
        pickle.dump((rels, adj, node_ids), fh, protocol=2)
    raw = []
    for s in case.batches:
        b = case.batches[s]
        for i in range(len(b["target"])):
            qg = SynthKG.query_graph(s, b["rels"], b["target"][i], b["anchors"][:, i])
            negs = [int(x) for x in b["negs"][i]][:n_neg_keep(i)]
            raw.append((qg, negs, negs[:1] if "inter" in s else None))
    with open(tmp_path / "queries.pkl", "wb") as fh:
        pickle.dump(raw, fh, protocol=2)
    return raw


@pytest.fixture()
def dataset(tmp_path):
    case = make_case(seed=9, d=32, decoder="bilinear", inter="mean", n_queries=23, n_neg=5)
    raw = _write_dataset(tmp_path, case, lambda i: 1 if i % 3 == 0 else 5)
    return tmp_path, case, raw


def test_load_graph_tables_and_row_lookup(dataset):
    path, case, _ = dataset
    torch.manual_seed(0)
    graph, feature_modules, node_maps = data.load_graph(str(path), 16)
    kg = case.kg
    assert set(feature_modules) == set(kg.modes)
    for m in kg.modes:
        assert feature_modules[m].weight.shape == (kg.sizes[m] + 2, 16)       # N + 2 rows (bio/data_utils.py:14-17)
        assert node_maps[m][-1] == -1
        ids = kg.node_ids[m][:50]
        want = np.array([node_maps[m][int(n)] + 1 for n in ids])              # bio/data_utils.py:21
        assert np.array_equal(graph.features(ids, m), want)
        assert graph.features([-1], m).tolist() == [0]
        assert abs(float(feature_modules[m].weight.std()) - 1.0 / 16) < 0.01
    assert graph.relations == kg.relations
    for m in kg.modes:                                                          # graph.py:116-120
        want = set()
        for r in kg.rel_keys:
            if r[0] == m:
                want |= set(int(n) for n in kg.node_ids[m][:40])
        assert set(graph.full_lists[m]) == want
    # the operator stack is constructible from what load_graph returns, like bio/train.py:51-55
    enc = gqe.get_encoder(0, graph, graph.feature_dims, feature_modules, cuda=False)
    assert enc.rows(kg.node_ids[kg.modes[0]][:4], kg.modes[0]).dtype == np.int32


def test_query_loaders_group_like_the_reference(dataset):
    path, case, raw = dataset
    f = str(path / "queries.pkl")
    flat = data.load_queries(f)
    assert len(flat) == len(raw)
    for q, r in zip(flat, raw):
        # graph.py:59-62,98-100: deserialize passes len(negs) as the cap, and a list AT the cap goes
        # through random.sample -- a permutation of the stored negatives
        assert sorted(q.neg_samples) == sorted(r[1]) and q.hard_neg_samples == r[2]
    by_formula = data.load_queries_by_formula(f)
    assert set(by_formula) == set(case.batches)
    for s in case.batches:
        (formula, qs), = by_formula[s].items()
        assert formula == case.formula(s, cls=gqe.Formula)
        assert [q.target_node for q in qs] == [int(x) for x in case.batches[s]["target"]]
    by_type = data.load_queries_by_type(f)
    assert {k: len(v) for k, v in by_type.items()} == {s: 23 for s in case.batches}
    assert by_type["2-chain"][0].query_graph is not None                        # keep_graph=True default
    test = data.load_test_queries_by_formula(f)
    for s in case.batches:
        n_one = sum(len(v) for v in test["one_neg"][s].values())
        n_full = sum(len(v) for v in test["full_neg"][s].values())
        assert n_one == len([i for i in range(23) if i % 3 == 0]) and n_one + n_full == 23


class _Capture(object):
    def __init__(self):
        self.calls = []

    def margin_loss(self, formula, queries, hard_negatives=False):
        self.calls.append((formula, [q.target_node for q in queries], hard_negatives))
        return len(queries)


class _KeysList(dict):          # py2 semantics of dict.keys() for train_helpers.py:100
    def keys(self):
        return list(dict.keys(self))


def test_pick_batch_slicing_wraps_like_train_helpers():
    class Q(object):
        def __init__(self, t):
            self.target_node = t
    tq = {"f": [Q(i) for i in range(10)]}
    got = [[q.target_node for q in data.pick_batch(tq, it, 4)[1]] for it in range(4)]
    # start = (it*4) % 10, end = ((it+1)*4) % 10, end = n when it wrapped (train_helpers.py:101-104)
    assert got == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9], [2, 3, 4, 5]]


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_run_batch_matches_reference_source(dataset):
    path, case, _ = dataset
    ref_run_batch = ref_shim.reference_run_batch()
    by_formula = data.load_queries_by_formula(str(path / "queries.pkl"))
    # a multi-formula pool: every structure's formula in one dict, as train_queries[query_type] can hold
    pool = _KeysList()
    for s in by_formula:
        for f, qs in by_formula[s].items():
            pool[f] = qs
    for it in range(12):
        a, b = _Capture(), _Capture()
        np.random.seed(100 + it)
        r1 = ref_run_batch(pool, a, it, 7, hard_negatives=bool(it % 2))
        np.random.seed(100 + it)
        r2 = data.run_batch(pool, b, it, 7, hard_negatives=bool(it % 2))
        assert r1 == r2 and a.calls == b.calls


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_deserialised_queries_match_reference_query_class(dataset):
    path, case, raw = dataset
    g = ref_shim.load()[0]
    random.seed(21)
    loaded = data.load_queries(str(path / "queries.pkl"))
    random.seed(21)                      # deserialize draws from the global random stream (graph.py:62)
    for info, mine in zip(raw, loaded):
        theirs = g.Query.deserialize(info)
        assert mine.anchor_nodes == theirs.anchor_nodes and mine.target_node == theirs.target_node
        assert mine.formula.query_type == theirs.formula.query_type and mine.formula.rels == theirs.formula.rels
        assert mine.neg_samples == theirs.neg_samples and mine.hard_neg_samples == theirs.hard_neg_samples
