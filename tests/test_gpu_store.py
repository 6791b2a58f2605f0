"""Device-resident query store (store.DeviceQueryStore, gqe_margin_loss_store_device): batches sliced and
negatives drawn on the GPU, then the ordinary fused scoring -- checked against the host path on the pairs the
GPU actually drew (netquery/train_helpers.py:95-107, netquery/model.py:112-127)."""
import numpy as np
import pytest
import torch

import graphqembed_b200 as gqe
from graphqembed_b200.store import DeviceSlice, QueryStore
from graphqembed_b200.synth import SynthKG
from helpers import build_package_model
from oracle.cases import make_case

pytestmark = pytest.mark.gpu


def _records(case, n_keep=lambda i: 5):
    raw = []
    for s in case.batches:
        b = case.batches[s]
        for i in range(len(b["target"])):
            qg = SynthKG.query_graph(s, b["rels"], b["target"][i], b["anchors"][:, i])
            negs = [int(x) for x in b["negs"][i]][:n_keep(i)]
            raw.append((qg, negs, negs[:2] if "inter" in s else None))
    return raw


@pytest.fixture(scope="module", params=[("bilinear", "mean", 128), ("transe", "min-simple", 64)], ids=["bilinear-d128", "transe-d64"])
def setup(request):
    decoder, inter, d = request.param
    case = make_case(seed=12, d=d, decoder=decoder, inter=inter, n_queries=700, n_neg=6, nodes_per_mode=900)
    model = build_package_model(case)
    store = QueryStore.from_records(_records(case, lambda i: 1 + i % 6))
    return case, model, store, store.to_device(model.device)


def _hinge(scores, margin=1.0):
    s = scores.double()
    return torch.clamp(margin - (s[:, 0] - s[:, 1]), min=0).mean().item()


def test_device_slices_score_the_pairs_they_draw(setup):
    case, model, store, dstore = setup
    model.negative_seed = 1000
    with torch.no_grad():
        for f in store.formulas():
            blk, dblk = store[f], dstore[f]
            start, stop = 37, len(blk) - 11
            sl = dblk.window(start, stop)
            assert isinstance(sl, DeviceSlice) and len(sl) == stop - start
            loss, pairs, scores = model._margin_loss_store([(f, sl)], return_pairs=True, return_scores=True)
            pairs_h = pairs.cpu().numpy()
            np.testing.assert_array_equal(pairs_h[:, 0], blk.targets[start:stop])
            if f.query_type == "1-chain":       # any node of the target mode (model.py:118-119)
                assert set(pairs_h[:, 1].tolist()) <= set(model._full_array(f.target_mode).tolist())
            else:
                for i in range(len(sl)):
                    q = start + i
                    assert pairs_h[i, 1] in blk.negs[blk.neg_ptr[q]:blk.neg_ptr[q + 1]]
            # the same pairs through the host-array path: same kernels, same bits
            want = model.score_batch(gqe.QueryBatch(f, blk.anchors[:, start:stop], pairs_h.reshape(-1))).reshape(-1, 2)
            assert torch.equal(scores, want)
            assert abs(loss.item() - _hinge(scores)) < 1e-6


def test_draw_is_a_function_of_the_seed_and_uniform_over_the_lists(setup):
    case, model, store, dstore = setup
    f = next(f for f in store.formulas() if f.query_type == "3-inter")
    blk, sl = store[f], dstore[f].all()
    with torch.no_grad():
        model.negative_seed, model._store_calls = 77, 0
        _, p1, _ = model._margin_loss_store([(f, sl)], return_pairs=True)
        model.negative_seed, model._store_calls = 77, 0
        _, p2, _ = model._margin_loss_store([(f, sl)], return_pairs=True)
        _, p3, _ = model._margin_loss_store([(f, sl)], return_pairs=True)      # the next call: seed + 1
        assert torch.equal(p1, p2)
        assert not torch.equal(p1, p3)
        # index work is held to the bit: the host restatement of the generator picks the same negatives
        from graphqembed_b200.store import device_draw
        n = len(blk)
        pick = device_draw(77, np.arange(n), np.diff(blk.neg_ptr))
        np.testing.assert_array_equal(p1.cpu().numpy()[:, 1], blk.negs[blk.neg_ptr[:-1] + pick])
        pick3 = device_draw(78, np.arange(n), np.diff(blk.neg_ptr))
        np.testing.assert_array_equal(p3.cpu().numpy()[:, 1], blk.negs[blk.neg_ptr[:-1] + pick3])
        # every position of a 6-long list is drawn about equally often over many seeds
        lens = np.diff(blk.neg_ptr)
        six = np.array([q for q in np.nonzero(lens == 6)[0] if len(set(blk.negs[blk.neg_ptr[q]:blk.neg_ptr[q] + 6].tolist())) == 6])
        lists = np.stack([blk.negs[blk.neg_ptr[q]:blk.neg_ptr[q] + 6] for q in six])
        counts = np.zeros(6)
        for _ in range(40):
            _, p, _ = model._margin_loss_store([(f, sl)], return_pairs=True)
            drawn = p.cpu().numpy()[six, 1]
            counts += np.bincount(np.argmax(lists == drawn[:, None], axis=1), minlength=6)
        share = counts / counts.sum()
        assert counts.sum() > 2000 and share.min() > 0.13 and share.max() < 0.20, share


def test_mix_of_device_slices_is_one_call(setup):
    case, model, store, dstore = setup
    items = [(f, dstore[f].window(5, 5 + 100 + 13 * i)) for i, f in enumerate(store.formulas())]
    with torch.no_grad():
        ctx = model.context()
        l0 = ctx.launch_count()
        loss = model.margin_loss_mix(items)
        launches = ctx.launch_count() - l0
        model.negative_seed, model._store_calls = 5, 0
        loss2, pairs, scores = model._margin_loss_store(items, return_pairs=True, return_scores=True)
    assert launches <= 4          # the batch kernel + the fused kernel (+ weight preparation on first use)
    assert abs(loss2.item() - _hinge(scores)) < 1e-6
    assert 0.0 < loss.item() < 2.0
    n = sum(len(sl) for _, sl in items)
    assert tuple(pairs.shape) == (n, 2)
    q0 = 0
    for f, sl in items:            # every slice landed at its own offset
        np.testing.assert_array_equal(pairs[q0:q0 + len(sl), 0].cpu().numpy(), store[f].targets[sl.start:sl.stop])
        q0 += len(sl)
    # public entry point on ONE slice = the mix of one
    f, sl = items[2]
    with torch.no_grad():
        assert 0.0 < model.margin_loss(f, sl).item() < 2.0


def test_hard_negatives_and_errors(setup):
    case, model, store, dstore = setup
    chain = next(f for f in store.formulas() if f.query_type == "2-chain")
    inter = next(f for f in store.formulas() if f.query_type == "2-inter")
    with torch.no_grad():
        with pytest.raises(Exception, match="Hard negative"):
            model.margin_loss(chain, dstore[chain].window(0, 10), hard_negatives=True)
        blk = store[inter]
        _, pairs, _ = model._margin_loss_store([(inter, dstore[inter].window(0, 50))], hard_negatives=True, return_pairs=True)
        ph = pairs.cpu().numpy()
        for i in range(50):
            assert ph[i, 1] in blk.hards[blk.hard_ptr[i]:blk.hard_ptr[i + 1]]
    # a query without negatives: the reference's random.choice raises IndexError
    raw = _records(case, lambda i: 0 if i == 3 else 2)
    bad = QueryStore.from_records(raw)
    dbad = bad.to_device(model.device)
    with torch.no_grad():
        with pytest.raises(IndexError, match="no negative"):
            model.margin_loss(inter, dbad[inter].window(0, 20))
        assert 0.0 < model.margin_loss(inter, dbad[inter].window(4, 20)).item() < 2.0     # the context recovered
    with pytest.raises(IndexError):
        dstore[inter].window(0, len(store[inter]) + 1)


def test_training_falls_back_to_the_host_arrays(setup):
    case, model, store, dstore = setup
    f = next(f for f in store.formulas() if f.query_type == "2-chain")
    try:
        loss = model.margin_loss(f, dstore[f].window(0, 64))
        assert loss.requires_grad
        loss.backward()
        assert any(p.grad is not None for p in model.parameters())
    finally:
        model.zero_grad()
