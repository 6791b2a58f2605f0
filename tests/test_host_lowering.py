"""Host lowering produces the reference's indices, bit for bit (no GPU needed)."""
import numpy as np
import pytest

import graphqembed_b200 as gqe
from graphqembed_b200 import _lib
from graphqembed_b200.synth import N_ANCHORS, STRUCTURES, SynthKG, bio_shaped
from helpers import GOLDEN_FILES, GOLDEN_IDS, load_golden
from oracle import netquery_oracle as O


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=GOLDEN_IDS)
def test_rows_and_relation_order_match_reference_trace(path):
    case, exp = load_golden(path)
    lookup = gqe.RowLookup(case.kg.node_maps())
    for s in case.batches:
        f = case.formula(s, cls=gqe.Formula)
        qs = case.queries(s, cls=gqe.Query)
        batch, order = gqe.QueryBatch.from_queries(f, qs, [q.target_node for q in qs])
        assert order is None
        trace = exp["trace"][s]
        ref_rows = [t for t in trace if t[0] == "rows"]
        ref_rels = [tuple(t[1]) for t in trace if t[0] == "rel"]
        # the reference looks the target up first, then anchors 0..A-1 (model.py:72-92)
        assert ref_rows[0][1] == f.target_mode
        np.testing.assert_array_equal(lookup.rows(batch.targets, f.target_mode), np.array(ref_rows[0][2], dtype=np.int32))
        for k, mode in enumerate(f.anchor_modes):
            assert ref_rows[1 + k][1] == mode
            np.testing.assert_array_equal(lookup.rows(batch.anchors[k], mode), np.array(ref_rows[1 + k][2], dtype=np.int32))
        assert gqe.relation_order(f) == ref_rels


def test_plan_fields():
    kg = bio_shaped(seed=1, scale=0.001)
    mode_ids = {m: i for i, m in enumerate(kg.modes)}
    rel_ids = {r: i for i, r in enumerate(kg.rel_keys)}
    rng = np.random.RandomState(0)
    for s in STRUCTURES:
        rels = kg.sample_rels(s, rng)
        f = gqe.Formula(s, rels)
        plan = gqe.lower_formula(f, mode_ids, rel_ids)
        assert plan.structure == _lib.STRUCTURE_ID[s]
        assert plan.target_mode == mode_ids[rels[0][0]]
        assert [m for m in plan.anchor_mode if m >= 0] == [mode_ids[m] for m in f.anchor_modes]
        assert len(f.anchor_modes) == N_ANCHORS[s]
        order = gqe.relation_order(f)
        assert [r for r in plan.rel if r >= 0] == [rel_ids[r] for r in order]
        if s == "3-chain_inter":
            assert plan.inter_mode == mode_ids[rels[0][-1]]
        elif "inter" in s:
            assert plan.inter_mode == plan.target_mode
        else:
            assert plan.inter_mode == -1
    assert len(kg.rel_keys) == 42 and sum(kg.sizes.values()) >= 5 * 8


def test_reversed_relations_are_distinct_parameters():
    f = gqe.Formula("2-inter", (("a", "r", "b"), ("a", "s", "c")))
    assert gqe.relation_order(f) == [("b", "r", "a"), ("c", "s", "a")]
    f = gqe.Formula("3-inter_chain", (("a", "r", "b"), (("a", "s", "c"), ("c", "t", "d"))))
    assert gqe.relation_order(f) == [("b", "r", "a"), ("d", "t", "c"), ("c", "s", "a")]
    assert f.anchor_modes == ("b", "d")
    f = gqe.Formula("3-chain_inter", (("a", "r", "b"), (("b", "s", "c"), ("b", "t", "d"))))
    assert gqe.relation_order(f) == [("c", "s", "b"), ("d", "t", "b"), ("b", "r", "a")]
    assert f.anchor_modes == ("c", "d")


def test_row_lookup_variants():
    ids = {"m": np.array([50, 10, 30], dtype=np.int64)}
    as_dict = {"m": {50: 0, 10: 1, 30: 2, -1: -1}}       # bio/data_utils.py:14-15 adds -1 -> -1
    for lk in (gqe.RowLookup(ids), gqe.RowLookup(as_dict)):
        np.testing.assert_array_equal(lk.rows([30, 50, 10, 10], "m"), [3, 1, 2, 2])
        with pytest.raises(KeyError):
            lk.rows([11], "m")
        with pytest.raises(KeyError):
            lk.rows([10], "nope")
    np.testing.assert_array_equal(gqe.RowLookup(as_dict).rows([-1], "m"), [0])
    np.testing.assert_array_equal(gqe.RowLookup(None).rows([0, 7], "m"), [1, 8])
    assert gqe.RowLookup(ids).rows([], "m").shape == (0,)


def test_from_queries_groups_repeated_queries_like_eval():
    """utils.py:86-88 call shape: batch + each query repeated per negative."""
    case, _ = load_golden(GOLDEN_FILES[0])
    s = "3-inter"
    f = case.formula(s, cls=gqe.Formula)
    qs = case.queries(s, cls=gqe.Query)[:5]
    rep = [q for q in qs for _ in q.neg_samples]
    targets = [q.target_node for q in qs] + [n for q in qs for n in q.neg_samples]
    batch, order = gqe.QueryBatch.from_queries(f, qs + rep, targets)
    assert batch.n_queries == 5 and batch.n_pairs == len(targets)
    k = len(qs[0].neg_samples)
    np.testing.assert_array_equal(batch.offsets, np.arange(6) * (k + 1))
    for i, q in enumerate(qs):
        assert tuple(batch.anchors[:, i]) == q.anchor_nodes
        mine = batch.targets[batch.offsets[i]:batch.offsets[i + 1]]
        assert list(mine) == [q.target_node] + list(q.neg_samples)
    # order maps batch order back to call order
    np.testing.assert_array_equal(np.asarray(targets)[order], batch.targets)


def test_query_batch_validation():
    f = gqe.Formula("2-inter", (("a", "r", "b"), ("a", "s", "c")))
    with pytest.raises(ValueError):
        gqe.QueryBatch(f, np.zeros((1, 4)), np.zeros(4))
    with pytest.raises(ValueError):
        gqe.QueryBatch(f, np.zeros((2, 4)), np.zeros(6))
    with pytest.raises(ValueError):
        gqe.QueryBatch(f, np.zeros((2, 2)), np.zeros(3), offsets=[0, 1, 2])
    b = gqe.QueryBatch(f, np.zeros((2, 0)), np.zeros(0))
    assert b.n_queries == 0 and b.n_pairs == 0


def test_synthetic_graph_is_closed_under_reversal():
    kg = bio_shaped(seed=0, scale=0.001)
    keys = set(kg.rel_keys)
    for r in kg.rel_keys:
        assert gqe.reverse_relation(r) in keys
    rng = np.random.RandomState(1)
    for s in STRUCTURES:
        for _ in range(20):
            rels = kg.sample_rels(s, rng)
            for r in gqe.relation_order(gqe.Formula(s, rels)):
                assert r in keys


def test_eval_batch_layout_is_positive_then_negatives_per_query():
    """evaluation._batch_arrays: one flat ragged batch per eval batch (utils.py:78-88 without
    the K-fold repetition of the query objects)."""
    from graphqembed_b200.evaluation import _batch_arrays
    import graphqembed_b200 as gqe

    class Q(object):
        def __init__(self, t, a):
            self.target_node, self.anchor_nodes = t, a
    f = gqe.Formula("2-inter", (("a", "r0", "b"), ("a", "r1", "c")))
    qs = [Q(10, (1, 2)), Q(11, (3, 4)), Q(12, (5, 6))]
    batch, offsets = _batch_arrays(f, qs, [100, 101, 200, 300, 301, 302], [2, 1, 3])
    assert offsets.tolist() == [0, 3, 5, 9]
    assert batch.targets.tolist() == [10, 100, 101, 11, 200, 12, 300, 301, 302]
    assert batch.anchors.tolist() == [[1, 3, 5], [2, 4, 6]]
    assert batch.offsets.tolist() == offsets.tolist()
    # equal lengths -> the regular layout (no offsets), which the tensor-core path takes
    batch, offsets = _batch_arrays(f, qs, [100, 200, 300], [1, 1, 1])
    assert batch.offsets is None and batch.targets.tolist() == [10, 100, 11, 200, 12, 300]
    # a query without negatives keeps its positive
    batch, offsets = _batch_arrays(f, qs, [100], [0, 1, 0])
    assert batch.targets.tolist() == [10, 11, 100, 12] and offsets.tolist() == [0, 1, 3, 4]
