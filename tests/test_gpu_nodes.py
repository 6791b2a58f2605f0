"""GPU: node ids lowered to rows INSIDE the kernels (gqe_bind_node_maps + the *_nodes entry
points), index bounds checks, the packed-weight cache and the streaming kernel of the
contraction-free decoders -- all through the C ABI, checked against the oracle / the row path."""
import copy
import pickle
import random

import numpy as np
import pytest
import torch

import graphqembed_b200 as gqe
from graphqembed_b200 import _lib
from graphqembed_b200.store import QueryStore
from graphqembed_b200.synth import STRUCTURES, SynthKG
from helpers import build_package_model, query_batch
from oracle.cases import make_case

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module", params=[(128, "bilinear", "mean"), (256, "bilinear", "min"), (64, "transe", "min-simple")],
                ids=lambda p: "d%d-%s-%s" % p)
def setup(request):
    d, decoder, inter = request.param
    case = make_case(seed=11 + d, d=d, decoder=decoder, inter=inter, n_queries=300, n_neg=1, nodes_per_mode=700)
    return case, build_package_model(case)


def test_node_ids_on_device_equal_host_lowered_rows(setup):
    """row = node_maps[mode][n] + 1 (bio/data_utils.py:21) done by the kernel == done on the host."""
    case, model = setup
    ctx = model.context()
    assert model._state[4], "node maps were not bound"
    for s in case.batches:
        b = case.batches[s]
        pairs = np.stack([b["target"], b["negs"][:, 0]], 1)
        batch = query_batch(case, s, pairs)
        plan = model.plan(batch.formula)
        rows_a, rows_t = model.lower_batch(batch)
        ids_a, ids_t = batch.int32_ids()
        out_r = np.empty(batch.n_pairs, np.float32)
        out_n = np.empty(batch.n_pairs, np.float32)
        ctx.score_host(plan, rows_a, rows_t, None, out_r)
        ctx.score_host(plan, ids_a, ids_t, None, out_n, nodes=True)
        np.testing.assert_array_equal(out_n, out_r, err_msg=s)
        np.testing.assert_array_equal(_np(model.score_batch(batch)), out_r, err_msg=s)     # device entry point, node ids
        lr, ln = np.empty(1, np.float32), np.empty(1, np.float32)
        ctx.margin_loss_host(plan, rows_a, rows_t, 1.0, lr)
        ctx.margin_loss_host(plan, ids_a, ids_t, 1.0, ln, nodes=True)
        assert lr[0] == ln[0] == model.margin_loss_batch(batch).item()
    # grouped, all structures in one call
    batches = [query_batch(case, s, np.stack([case.batches[s]["target"], case.batches[s]["negs"][:, 0]], 1))
               for s in case.batches]
    loss_nodes = model.margin_loss_grouped(batches).item()
    st = model._state
    st[4] = False                                   # force the host-lowered row path
    loss_rows = model.margin_loss_grouped(batches).item()
    st[4] = True
    assert loss_nodes == loss_rows


def test_unknown_node_raises_keyerror_like_the_reference(setup):
    case, model = setup
    ctx = model.context()
    s = "2-inter"
    b = case.batches[s]
    pairs = np.stack([b["target"], b["negs"][:, 0]], 1)
    good = query_batch(case, s, pairs)
    plan = model.plan(good.formula)
    ids_a, ids_t = [x.copy() for x in good.int32_ids()]
    ids_a[1, 17] = 10 ** 8                          # not a node of any mode
    out = np.empty(good.n_pairs, np.float32)
    with pytest.raises(KeyError) as e:
        ctx.score_host(plan, ids_a, ids_t, None, out, nodes=True)
    assert "100000000" in str(e.value) and isinstance(e.value, gqe.GqeIndexError)
    ids_t2 = ids_t.copy()
    ids_t2[5] = -12345
    loss = np.empty(1, np.float32)
    with pytest.raises(KeyError):
        ctx.margin_loss_host(plan, good.int32_ids()[0], ids_t2, 1.0, loss, nodes=True)
    # the error is consumed: a clean call afterwards succeeds
    ctx.score_host(plan, good.int32_ids()[0], ids_t, None, out, nodes=True)
    # through the model (device entry point): raised inside the call, like the reference's forward
    bad = gqe.QueryBatch(good.formula, ids_a, ids_t)
    with pytest.raises(KeyError):
        model.score_batch(bad)
    assert np.isfinite(_np(model.score_batch(good))).all()
    # asynchronous use: the error waits in the context until polled
    model.check_indices = False
    model.score_batch(bad)
    with pytest.raises(KeyError):
        model.context().index_error()
    model.context().index_error()                   # cleared
    model.check_indices = True


def test_row_index_out_of_range_is_reported_not_read(setup):
    """rows were never checked against the bound table sizes."""
    case, model = setup
    ctx = model.context()
    s = "3-chain"
    b = case.batches[s]
    batch = query_batch(case, s, np.stack([b["target"], b["negs"][:, 0]], 1))
    plan = model.plan(batch.formula)
    rows_a, rows_t = model.lower_batch(batch)
    out = np.empty(batch.n_pairs, np.float32)
    for bad_value in (10 ** 7, -3):
        t = rows_t.copy()
        t[11] = bad_value
        with pytest.raises(IndexError) as e:
            ctx.score_host(plan, rows_a, t, None, out)
        assert str(bad_value) in str(e.value)
    # operator-level encode and its backward
    enc = model.enc
    mode = case.kg.modes[0]
    rows = torch.tensor([1, 2, 10 ** 6], dtype=torch.int32, device="cuda")
    o = torch.empty((case.d, 3), device="cuda")
    c2 = enc._ctx()
    c2.encode_device(enc.mode_ids[mode], 3, rows.data_ptr(), o.data_ptr())
    with pytest.raises(IndexError):
        c2.index_error()
    g = torch.zeros_like(enc.table(mode))
    c2.encode_bwd_device(enc.mode_ids[mode], 3, rows.data_ptr(), o.data_ptr(), g.data_ptr())
    with pytest.raises(IndexError):
        c2.index_error()
    assert float(g[0].abs().sum()) == 0.0           # nothing was scattered for the bad row


def test_weight_cache_follows_in_place_updates(setup):
    case, model = setup
    if case.decoder != "bilinear":
        pytest.skip("the weight cache belongs to the tensor-core path")
    s = "3-inter_chain"
    b = case.batches[s]
    batch = query_batch(case, s, np.stack([b["target"], b["negs"][:, 0]], 1))
    ctx = model.context()
    model.margin_loss_batch(batch)
    preps, launches = ctx.weight_prep_count(), ctx.launch_count()
    l1 = model.margin_loss_batch(batch).item()
    assert ctx.weight_prep_count() == preps, "matrices were re-packed although nothing changed"
    assert ctx.launch_count() - launches == 1, "a cached call must be the fused kernel alone"
    # an optimiser-style in-place update is seen (autograd version counter -> gqe_invalidate_weights)
    rel = batch.formula.rels[0]
    key = gqe.reverse_relation(rel)
    with torch.no_grad():
        model.path_dec.mats[key].mul_(1.5)
        model.inter_dec.pre_mats[batch.formula.target_mode].add_(0.01)
    l2 = model.margin_loss_batch(batch).item()
    assert ctx.weight_prep_count() > preps
    fresh = copy.deepcopy(model)                    # a new context, nothing cached
    assert fresh._state is None
    assert fresh.margin_loss_batch(batch).item() == l2 != l1
    # raw C ABI contract: stale until invalidated; cache off = always live
    with torch.no_grad():
        model.path_dec.mats[key].data.add_(0.03)    # .data: no version bump -> the model cannot see it
    stale = model.margin_loss_batch(batch).item()
    assert stale == l2
    model.context().invalidate_weights()
    l3 = model.margin_loss_batch(batch).item()
    assert l3 != l2
    ctx.set_weight_cache(False)
    with torch.no_grad():
        model.path_dec.mats[key].data.sub_(0.03)
    # (with the cache off small batches are not pre-composed: same algebra, different rounding)
    assert abs(model.margin_loss_batch(batch).item() - l2) < 2e-6
    ctx.set_weight_cache(True)


def test_model_survives_deepcopy_and_pickle(setup):
    """a ctypes context inside the module broke copy.deepcopy / torch.save(model)."""
    case, model = setup
    s = "2-chain"
    b = case.batches[s]
    batch = query_batch(case, s, np.stack([b["target"], b["negs"][:, 0]], 1))
    want = model.margin_loss_batch(batch).item()
    e = model.enc.forward([int(x) for x in b["target"][:8]], batch.formula.target_mode)   # creates the operator's own context
    clone = pickle.loads(pickle.dumps(model))
    assert clone.margin_loss_batch(batch).item() == want
    assert torch.equal(clone.enc.forward([int(x) for x in b["target"][:8]], batch.formula.target_mode), e)
    assert copy.deepcopy(model.path_dec).project(e, batch.formula.rels[0]).shape == e.shape


@pytest.mark.parametrize("decoder", ["transe", "bilinear-diag"])
@pytest.mark.parametrize("inter", ["mean-simple", "min-simple", "mean"])
@pytest.mark.parametrize("d", [32, 128, 256])
def test_streaming_kernel_vs_oracle_and_tile_kernel(decoder, inter, d):
    """gqe_fused_vec (T <= 2, no contraction) against the oracle, and against the tile kernel
    (T = 3 routes there) on the same pairs."""
    nq = 333
    case = make_case(seed=d + len(decoder), d=d, decoder=decoder, inter=inter, n_queries=nq, n_neg=2, nodes_per_mode=400)
    model = build_package_model(case)
    orc = case.oracle()
    for s in case.batches:
        if inter == "mean" and "inter" in s:
            continue                                 # DeepSets on a vector decoder stays on the tile kernel
        b = case.batches[s]
        t3 = np.concatenate([b["target"][:, None], b["negs"]], axis=1)                    # [Q, 3]
        qs, f = case.queries(s), case.formula(s)
        want = np.stack([orc.forward(f, qs, [int(x) for x in t3[:, j]]).numpy() for j in range(3)], 1)
        launches = model.context().launch_count()
        loss, sc = model.margin_loss_batch(query_batch(case, s, t3[:, :2]), margin=1, return_scores=True)
        assert model.context().launch_count() - launches == 1
        np.testing.assert_allclose(_np(sc), want[:, :2], rtol=0, atol=2e-5, err_msg=s)
        assert abs(loss.item() - np.maximum(0, 1 - (want[:, 0] - want[:, 1])).mean()) <= 2e-5
        tile = _np(model.score_batch(query_batch(case, s, t3))).reshape(nq, 3)           # gqe_fused_simt
        np.testing.assert_allclose(_np(sc), tile[:, :2], rtol=0, atol=2e-6, err_msg=s)
        single = _np(model.score_batch(query_batch(case, s, t3[:, 0])))                   # T = 1
        np.testing.assert_allclose(single, tile[:, 0], rtol=0, atol=2e-6, err_msg=s)
        assert model.margin_loss_batch(query_batch(case, s, t3[:, :2])).item() == loss.item()   # deterministic


def test_store_slice_is_scored_without_query_objects(setup):
    case, model = setup
    raw = []
    for s in case.batches:
        b = case.batches[s]
        for i in range(len(b["target"])):
            negs = [int(x) for x in b["negs"][i]]
            raw.append((SynthKG.query_graph(s, b["rels"], b["target"][i], b["anchors"][:, i]), negs,
                        negs if "inter" in s else None))
    store = QueryStore.from_records(raw)
    model.reference_negatives = True
    for s in case.batches:
        f = case.formula(s, cls=gqe.Formula)
        qs = [gqe.Query(r[0], r[1], r[2], len(r[1]) + 1) for r in raw if r[0][0] == s]
        sl = store[f].window(10, 250)
        for hard in ((False, True) if "inter" in s else (False,)):
            random.seed(3)
            want = model.margin_loss(f, qs[10:250], hard_negatives=hard).item()
            random.seed(3)
            got = model.margin_loss(f, sl, hard_negatives=hard).item()
            assert got == want, (s, hard)
        np.testing.assert_array_equal(_np(model.forward(f, sl, sl.targets)),
                                      _np(model.forward(f, qs[10:250], [q.target_node for q in qs[10:250]])))
    model.reference_negatives = False
    model.negative_rng = np.random.default_rng(0)
    f = case.formula("3-inter", cls=gqe.Formula)
    assert np.isfinite(model.margin_loss(f, store[f].all()).item())
    with pytest.raises(Exception):
        model.margin_loss(case.formula("2-chain", cls=gqe.Formula), store[case.formula("2-chain", cls=gqe.Formula)].all(),
                          hard_negatives=True)


def test_margin_loss_mix_equals_weighted_per_formula_losses():
    """margin_loss_mix([(formula, StoreSlice)]): one grouped launch, the same negatives (same Generator
    state) and the size-weighted mean of the per-formula losses."""
    from graphqembed_b200.store import QueryStore
    from graphqembed_b200.synth import SynthKG
    from helpers import build_package_model
    from oracle.cases import make_case
    case = make_case(seed=5, d=128, decoder="bilinear", inter="mean", n_queries=300, n_neg=4, nodes_per_mode=900)
    model = build_package_model(case)
    raw = []
    for s in case.batches:
        b = case.batches[s]
        for i in range(len(b["target"])):
            negs = [int(x) for x in b["negs"][i]]
            raw.append((SynthKG.query_graph(s, b["rels"], b["target"][i], b["anchors"][:, i]), negs,
                        negs if "inter" in s else None))
    store = QueryStore.from_records(raw)
    items = [(case.formula(s, cls=gqe.Formula), store[case.formula(s, cls=gqe.Formula)].window(0, 100 + 20 * k))
             for k, s in enumerate(case.batches)]
    with torch.no_grad():
        model.negative_rng = np.random.default_rng(42)
        want = sum(float(model.margin_loss(f, sl)) * len(sl) for f, sl in items) / sum(len(sl) for _, sl in items)
        model.negative_rng = np.random.default_rng(42)
        got = float(model.margin_loss_mix(items))
    assert abs(got - want) < 2e-6, (got, want)
