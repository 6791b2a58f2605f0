"""The oracle reproduces the frozen outputs of the real reference (tests/golden)."""
import random

import numpy as np
import pytest
import torch

from helpers import GOLDEN_FILES, GOLDEN_IDS, LOSS_SEED, load_golden
from oracle import netquery_oracle as O

# Frozen on one machine, replayed on another: sgemm summation order may differ
# between CPU micro-architectures, so the portable bound is a few fp32 ulps of
# a unit-scale cosine.  Bit-exactness is asserted in test_oracle_vs_reference.py
# on the machine that also runs the reference.
ATOL = 2e-6


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def test_golden_files_present():
    assert len(GOLDEN_FILES) >= 10


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=GOLDEN_IDS)
def test_oracle_matches_reference_outputs(path):
    case, exp = load_golden(path)
    orc = case.oracle()
    for s in case.batches:
        f = case.formula(s)
        qs = case.queries(s)
        b = case.batches[s]
        pos = orc.forward(f, qs, [q.target_node for q in qs]).numpy()
        neg = orc.forward(f, qs, [int(x) for x in b["negs"][:, 0]]).numpy()
        np.testing.assert_allclose(pos, exp[s + "/pos"], rtol=0, atol=ATOL)
        np.testing.assert_allclose(neg, exp[s + "/neg"], rtol=0, atol=ATOL)
        rep = [q for q in qs for _ in q.neg_samples]
        ev = orc.forward(f, qs + rep, [q.target_node for q in qs] + [n for q in qs for n in q.neg_samples]).numpy()
        np.testing.assert_allclose(ev, exp[s + "/eval"], rtol=0, atol=ATOL)
        random.seed(LOSS_SEED)
        loss = orc.margin_loss(f, qs).item()
        assert abs(loss - float(exp[s + "/loss"])) <= ATOL
        if "inter" in s:
            random.seed(LOSS_SEED)
            hard = orc.margin_loss(f, qs, hard_negatives=True).item()
            assert abs(hard - float(exp[s + "/hard"])) <= ATOL


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=GOLDEN_IDS)
def test_oracle_index_trace_is_bit_exact(path):
    """Same table rows, same relation keys, same order as the reference."""
    case, exp = load_golden(path)
    orc = case.oracle()
    for s in case.batches:
        qs = case.queries(s)
        orc.trace = []
        orc.forward(case.formula(s), qs, [q.target_node for q in qs])
        got = [list(t) if t[0] == "rows" else ["rel", list(t[1])] for t in orc.trace]
        assert got == exp["trace"][s]


@pytest.mark.parametrize("path", GOLDEN_FILES[:3], ids=GOLDEN_IDS[:3])
def test_float64_oracle_agrees(path):
    """fp32 reference outputs sit within fp32 rounding of the fp64 evaluation."""
    case, exp = load_golden(path)
    orc64 = case.oracle(dtype=torch.float64)
    for s in case.batches:
        qs = case.queries(s)
        pos = orc64.forward(case.formula(s), qs, [q.target_node for q in qs]).numpy()
        np.testing.assert_allclose(pos, exp[s + "/pos"], rtol=0, atol=2e-5)


def test_hard_negatives_rejected_for_chains():
    case, _ = load_golden(GOLDEN_FILES[0])
    with pytest.raises(Exception, match="Hard negative"):
        case.oracle().margin_loss(case.formula("2-chain"), case.queries("2-chain"), hard_negatives=True)


def test_zero_row_gives_nan_like_reference():
    case, _ = load_golden(GOLDEN_FILES[0])
    s = "2-inter"
    qs = case.queries(s)
    mode = case.formula(s).target_mode
    row = case.kg.node_maps()[mode][qs[0].target_node] + 1
    case.tables[mode][row].zero_()
    out = case.oracle().forward(case.formula(s), qs, [q.target_node for q in qs])
    assert torch.isnan(out[0]) and not torch.isnan(out[1:]).any() or (out != out).sum() >= 1
