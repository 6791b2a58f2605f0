"""Bit-exact pin of the oracle against the unmodified reference, when mounted.

Runs only where /root/reference exists (the build container).  On the GPU box
the same pin is carried by tests/golden/*.npz.
"""
import random

import pytest
import torch

from oracle import ref_shim
from oracle.cases import DECODERS, INTERS, make_case

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")


@pytest.mark.parametrize("decoder", DECODERS)
@pytest.mark.parametrize("inter", INTERS)
def test_bit_exact_scores_and_loss(decoder, inter):
    torch.set_num_threads(1)
    case = make_case(seed=100 + 10 * DECODERS.index(decoder) + INTERS.index(inter), d=64, decoder=decoder, inter=inter, n_queries=37, n_neg=4)
    ref, g = ref_shim.build_reference_model(case.tables, case.kg.node_maps(), case.kg.relations, case.rel_params,
                                            decoder, inter, case.pre, case.post, full_lists=case.kg.full_lists())
    orc = case.oracle()
    with torch.no_grad():
        for s in case.batches:
            rq, oq = case.queries(s, cls=g.Query), case.queries(s)
            rf, of = case.formula(s, cls=g.Formula), case.formula(s)
            targets = [q.target_node for q in oq]
            assert torch.equal(ref.forward(rf, rq, targets), orc.forward(of, oq, targets)), s
            random.seed(5)
            a = ref.margin_loss(rf, rq)
            random.seed(5)
            b = orc.margin_loss(of, oq)
            assert torch.equal(a, b), s
            if "inter" in s:
                random.seed(6)
                a = ref.margin_loss(rf, rq, hard_negatives=True)
                random.seed(6)
                b = orc.margin_loss(of, oq, hard_negatives=True)
                assert torch.equal(a, b), s


def test_query_and_formula_restatement_match_reference_classes():
    g = ref_shim.load()[0]
    case = make_case(seed=3, d=32, decoder="bilinear", inter="mean")
    import graphqembed_b200 as gqe
    for s in case.batches:
        for cls in (None, gqe.Query):
            mine = case.queries(s) if cls is None else case.queries(s, cls=cls)
            theirs = case.queries(s, cls=g.Query)
            for a, b in zip(mine, theirs):
                assert a.anchor_nodes == b.anchor_nodes and a.target_node == b.target_node
                assert a.formula.rels == b.formula.rels and a.formula.anchor_modes == b.formula.anchor_modes
                assert a.formula.target_mode == b.formula.target_mode
                assert a.neg_samples == b.neg_samples and a.hard_neg_samples == b.hard_neg_samples


def test_unknown_query_type_returns_none_like_reference():
    g = ref_shim.load()[0]
    case = make_case(seed=4, d=32, decoder="bilinear", inter="mean")
    ref, _ = ref_shim.build_reference_model(case.tables, case.kg.node_maps(), case.kg.relations, case.rel_params,
                                            "bilinear", "mean", case.pre, case.post)
    f = case.formula("2-chain", cls=g.Formula)
    f.query_type = "4-chain"
    assert ref.forward(f, [], []) is None
    of = case.formula("2-chain")
    of.query_type = "4-chain"
    assert case.oracle().forward(of, [], []) is None


def test_eval_restatement_matches_reference_eval_functions():
    """oracle eval_auc_queries / eval_perc_queries vs the reference's own utils.py:26-91
    (exec'd from source), each driving its own model on the same queries."""
    from oracle import netquery_oracle as O
    g = ref_shim.load()[0]
    ref_auc, ref_perc = ref_shim.reference_eval_functions()
    case = make_case(seed=11, d=32, decoder="bilinear", inter="mean", n_queries=57, n_neg=6)
    ref, _ = ref_shim.build_reference_model(case.tables, case.kg.node_maps(), case.kg.relations, case.rel_params,
                                            "bilinear", "mean", case.pre, case.post, full_lists=case.kg.full_lists())
    orc = case.oracle()
    structures = ("2-chain", "2-inter", "3-inter_chain")
    tq_ref = {case.formula(s, cls=g.Formula): case.queries(s, cls=g.Query) for s in structures}
    tq_orc = {case.formula(s): case.queries(s) for s in structures}
    with torch.no_grad():
        for hard in (False, True):
            a, fa = ref_auc(tq_ref, ref, batch_size=20, hard_negatives=hard, seed=3)
            b, fb = O.eval_auc_queries(tq_orc, orc, batch_size=20, hard_negatives=hard, seed=3)
            assert a == b
            assert [fa[f] for f in tq_ref] == [fb[f] for f in tq_orc]
            assert ref_perc(tq_ref, ref, batch_size=20, hard_negatives=hard) == \
                O.eval_perc_queries(tq_orc, orc, batch_size=20, hard_negatives=hard)
