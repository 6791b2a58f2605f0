"""QueryStore (graphqembed_b200/store.py): the flat form of the reference's query files, and the
O(1) node-id -> row tables of RowLookup -- host logic only, no GPU."""
import pickle
import random

import numpy as np
import pytest

import graphqembed_b200 as gqe
from graphqembed_b200 import data
from graphqembed_b200.lowering import RowLookup
from graphqembed_b200.store import QueryStore, batch_window
from graphqembed_b200.synth import SynthKG
from oracle.cases import make_case


def _records(case, n_keep=lambda i: 5):
    raw = []
    for s in case.batches:
        b = case.batches[s]
        for i in range(len(b["target"])):
            qg = SynthKG.query_graph(s, b["rels"], b["target"][i], b["anchors"][:, i])
            negs = [int(x) for x in b["negs"][i]][:n_keep(i)]
            raw.append((qg, negs, negs[:2] if "inter" in s else None))
    return raw


@pytest.fixture()
def case_and_records():
    case = make_case(seed=4, d=32, decoder="bilinear", inter="mean", n_queries=29, n_neg=6)
    return case, _records(case, lambda i: 1 if i % 4 == 0 else 6)


def test_store_blocks_hold_the_same_indices_as_query_objects(case_and_records, tmp_path):
    case, raw = case_and_records
    with open(tmp_path / "q.pkl", "wb") as fh:
        pickle.dump(raw, fh, protocol=2)
    store = QueryStore.from_file(str(tmp_path / "q.pkl"))
    by_formula = data.load_queries_by_formula(str(tmp_path / "q.pkl"))
    assert len(store) == len(raw)
    assert set(store.by_type) == set(by_formula)
    for qt in by_formula:
        assert store.formulas(qt) == list(by_formula[qt])
        for f, qs in by_formula[qt].items():
            blk = store[f]
            assert blk.anchors.dtype == np.int32 and blk.targets.dtype == np.int32
            assert blk.targets.tolist() == [q.target_node for q in qs]
            for k in range(len(f.anchor_modes)):
                assert blk.anchors[k].tolist() == [q.anchor_nodes[k] for q in qs]
            for i, q in enumerate(qs):     # deserialize permutes a list at its cap (graph.py:59-62): compare as sets
                assert sorted(blk.negs[blk.neg_ptr[i]:blk.neg_ptr[i + 1]].tolist()) == sorted(q.neg_samples)
                hard = blk.hards[blk.hard_ptr[i]:blk.hard_ptr[i + 1]].tolist()
                assert sorted(hard) == sorted(q.hard_neg_samples or [])
    split = QueryStore.test_split(str(tmp_path / "q.pkl"))
    ref = data.load_test_queries_by_formula(str(tmp_path / "q.pkl"))
    for name in ("full_neg", "one_neg"):
        assert len(split[name]) == sum(len(v) for t in ref[name].values() for v in t.values())


def test_batch_window_is_the_reference_window():
    # start = (it*B) % n, end = min(((it+1)*B) % n, n), end = n if end <= start (train_helpers.py:101-104)
    for n in (1, 7, 10, 512, 1000):
        for B in (1, 4, 512):
            for it in range(0, 40):
                start = (it * B) % n
                end = min(((it + 1) * B) % n, n)
                end = n if end <= start else end
                assert batch_window(it, B, n) == (start, end)


def test_sample_batch_draws_like_pick_batch(case_and_records):
    case, raw = case_and_records
    store = QueryStore.from_records(raw + raw[:40])
    # one pool holding every formula of one type, as train_queries[query_type] does
    qt = "2-inter"
    pool_lists = {f: [None] * len(store[f]) for f in store.formulas(qt)}
    for it in range(10):
        np.random.seed(it)
        f_ref, qs = data.pick_batch(pool_lists, it, 8)
        np.random.seed(it)
        sl = store.sample_batch(qt, it, 8)
        assert sl.formula == f_ref and len(sl) == len(qs)
        np.random.seed(it)
        f2, sl2 = data.pick_batch({f: store[f] for f in store.formulas(qt)}, it, 8)
        assert f2 == f_ref and (sl2.start, sl2.stop) == (sl.start, sl.stop)


def test_negative_draws(case_and_records):
    case, raw = case_and_records
    store = QueryStore.from_records(raw)
    f = store.formulas("3-inter")[0]
    sl = store[f].window(3, 20)
    qs = [gqe.Query(r[0], r[1], r[2], len(r[1]) + 1) for r in raw if r[0][0] == "3-inter"][3:20]
    for hard in (False, True):
        random.seed(5)
        want = [random.choice(q.hard_neg_samples if hard else q.neg_samples) for q in qs]     # model.py:116-120
        random.seed(5)
        got = sl.draw_negatives(hard=hard, reference_stream=True)
        assert got.tolist() == want and got.dtype == np.int32
        rng = np.random.default_rng(0)
        for _ in range(5):
            v = sl.draw_negatives(hard=hard, rng=rng)
            for x, q in zip(v.tolist(), qs):
                assert x in (q.hard_neg_samples if hard else q.neg_samples)
    full = np.arange(100, 140, dtype=np.int32)
    random.seed(9)
    want = [random.choice(list(full)) for _ in qs]
    random.seed(9)
    assert sl.draw_negatives(full_list=full, reference_stream=True).tolist() == [int(x) for x in want]
    chain = store[store.formulas("2-chain")[0]].all()
    with pytest.raises(IndexError):          # random.choice([]) raises IndexError in the reference
        chain.draw_negatives(hard=True)
    off, vals = sl.negative_lists()
    assert off[0] == 0 and off[-1] == len(vals) == sum(len(q.neg_samples) for q in qs)
    b = sl.margin_batch(sl.draw_negatives(rng=np.random.default_rng(1)))
    assert b.anchors.dtype == np.int32 and b.n_pairs == 2 * len(sl) and b.int32_ids()[0] is b.anchors


def test_row_lookup_dense_and_sparse_agree():
    rng = np.random.RandomState(0)
    dense_ids = rng.permutation(5000)[:3000] + 70
    sparse_ids = (rng.permutation(3000).astype(np.int64) * 10 ** 9) + 5
    L = RowLookup({"d": dense_ids, "s": sparse_ids, "dict": {int(n): i for i, n in enumerate(dense_ids[:50])}})
    assert "d" in L._lut and "s" in L._sorted
    for mode, ids in (("d", dense_ids), ("s", sparse_ids)):
        got = L.rows(ids[::7], mode)
        assert got.dtype == np.int32 and np.array_equal(got, np.arange(len(ids))[::7] + 1)
        assert np.array_equal(L.rows(ids[:6].reshape(2, 3), mode), (np.arange(6) + 1).reshape(2, 3))
        for bad in (int(ids.max()) + 1, int(ids.min()) - 1):
            with pytest.raises(KeyError):
                L.rows([int(ids[0]), bad], mode)
    with pytest.raises(KeyError):
        L.rows([71], "dict") if 71 not in dense_ids[:50] else (_ for _ in ()).throw(KeyError())
    with pytest.raises(KeyError):
        L.rows([1], "nope")
    assert RowLookup(None).rows([0, 4], "x").tolist() == [1, 5]
    assert L.device_maps(["d", "s"], [1, 1], "cpu") is None       # a sparse mode: host lookup only
    ptrs, bases, lens, keep = L.device_maps(["d"], [len(dense_ids) + 2], "cpu")
    assert bases == [int(dense_ids.min())] and lens == [int(dense_ids.max() - dense_ids.min() + 1)] and ptrs[0] != 0
    assert RowLookup(None).device_maps(["a", "b"], [10, 20], "cpu")[:3] == ([0, 0], [-1, -1], [10, 20])


def test_device_store_descriptors_on_cpu(case_and_records):
    """QueryStore.to_device (here onto the CPU device: the descriptor logic needs no GPU): same blocks, same
    batch sampling as the host store, slices that map back to the host views."""
    import torch
    from graphqembed_b200.store import DeviceQueryStore, DeviceSlice
    case, raw = case_and_records
    store = QueryStore.from_records(raw)
    dstore = store.to_device(torch.device("cpu"))
    assert isinstance(dstore, DeviceQueryStore) and len(dstore) == len(store)
    assert dstore.formulas() == store.formulas()
    for f in store.formulas():
        blk, dblk = store[f], dstore[f]
        assert len(dblk) == len(blk)
        np.testing.assert_array_equal(dblk.anchors.numpy(), blk.anchors)
        np.testing.assert_array_equal(dblk.targets.numpy(), blk.targets)
        np.testing.assert_array_equal(dblk.neg_ptr.numpy(), blk.neg_ptr)
        assert dblk.anchors.dtype == torch.int32 and dblk.neg_ptr.dtype == torch.int64
        sl = dblk.window(1, len(blk))
        assert isinstance(sl, DeviceSlice) and len(sl) == len(blk) - 1 and sl.formula == f
        np.testing.assert_array_equal(sl.host().targets, blk.targets[1:])
        with pytest.raises(IndexError):
            dblk.window(0, len(blk) + 1)
    qt = next(iter(store.by_type))
    np.random.seed(11)
    a = [store.sample_batch(qt, i, 4) for i in range(6)]
    np.random.seed(11)
    b = [dstore.sample_batch(qt, i, 4) for i in range(6)]
    for x, y in zip(a, b):
        assert x.formula == y.formula and (x.start, x.stop) == (y.start, y.stop)


def test_device_draw_restatement_is_uniform_and_in_range():
    from graphqembed_b200.store import device_draw
    lens = np.array([1, 2, 3, 7, 1000, 1 << 20] * 5000)
    pick = device_draw(123456789, np.arange(lens.size), lens)
    assert pick.dtype == np.int64 and (pick >= 0).all() and (pick < lens).all()
    sevens = pick[lens == 7]
    share = np.bincount(sevens, minlength=7) / sevens.size
    assert share.min() > 0.12 and share.max() < 0.17
    assert not np.array_equal(pick, device_draw(123456790, np.arange(lens.size), lens))
    # splitmix64 known answer: seed 0, position 0 -> first output of the generator seeded with 0
    z = device_draw(0, [0], [1 << 32])[0]
    assert z == 0xE220A8397B1DCDAF >> 32
