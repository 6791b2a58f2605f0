"""GPU: the sparse training step (row gradients + SparseRowAdam) follows the trajectory of the
reference's step -- dense table gradients + torch.optim.Adam (bio/train.py:59-62,
train_helpers.py:76-79) -- on every row of every table, touched or not."""
import copy
import random

import numpy as np
import pytest
import torch

import graphqembed_b200 as gqe
from helpers import build_package_model
from oracle.cases import make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("decoder,inter,d", [("bilinear", "mean", 128), ("transe", "min-simple", 64),
                                             ("bilinear-diag", "min", 32)])
def test_sparse_row_adam_matches_dense_adam(decoder, inter, d):
    case = make_case(seed=3, d=d, decoder=decoder, inter=inter, n_queries=160, n_neg=3, nodes_per_mode=400)
    dense_model = build_package_model(case)
    sparse_model = copy.deepcopy(dense_model)
    dense_opt = torch.optim.Adam(dense_model.parameters(), lr=0.01)
    sparse_opt = gqe.SparseRowAdam(sparse_model, lr=0.01)
    order = ["1-chain", "3-inter", "2-chain", "3-inter_chain", "1-chain", "2-inter", "3-chain_inter", "3-chain",
             "1-chain", "3-inter", "2-inter", "1-chain", "3-inter_chain", "2-chain"]
    for it, s in enumerate(order):
        f = case.formula(s, cls=gqe.Formula)
        qs = case.queries(s, cls=gqe.Query)
        lo = (it * 37) % 100
        batch = qs[lo:lo + 48]
        losses = []
        for model, opt in ((dense_model, dense_opt), (sparse_model, sparse_opt)):
            random.seed(1000 + it)
            opt.zero_grad()
            loss = model.margin_loss(f, batch, hard_negatives=("inter" in s and it % 2 == 1))
            loss.backward()
            opt.step()
            losses.append(loss.item())
        assert abs(losses[0] - losses[1]) <= 2e-5 * max(1.0, abs(losses[0])), (it, s, losses)
    # table gradients really were sparse, and only touched rows carry moments
    touched = 0
    for m in case.kg.modes:
        assert sparse_model.enc.table(m).grad is None or sparse_model.enc.table(m).grad.is_sparse
        touched += int((sparse_opt.state[m]["last"] > 0).sum())
    assert 0 < touched < sum(case.kg.sizes[m] + 2 for m in case.kg.modes)
    sparse_opt.flush()
    for m in case.kg.modes:
        a, b = dense_model.enc.table(m).detach(), sparse_model.enc.table(m).detach()
        assert torch.isfinite(b).all()
        diff = (a - b).abs().max(dim=1).values
        worst = torch.argsort(diff, descending=True)[:4]
        info = [(int(r), float(diff[r]), int(sparse_opt.state[m]["last"][r]), sparse_opt.state[m]["step"]) for r in worst]
        # Adam turns a gradient entry of ~1e-8 (cancellation noise: the dense path sums a row's
        # contributions with atomics, the sparse path in sorted order) into a step of ~lr/2, so a
        # handful of entries may sit anywhere within one step; everything else must agree closely
        # (the optimiser itself is held to 1e-5 in test_sparse_adam_equals_dense_adam_on_equal_gradients)
        close = torch.isclose(b, a, rtol=1e-3, atol=1e-5)
        assert close.float().mean().item() > 0.99, "%s %s" % (m, info)
        assert (a - b).abs().max().item() < 0.01 * len(order), "%s %s" % (m, info)
        moved = (a - case.tables[m].to(a.device)).abs().max().item()
        assert moved > 1e-3, "the tables did not train"
    for (na, pa), (nb, pb) in zip(dense_model.named_parameters(), sparse_model.named_parameters()):
        if "feat-" not in na:
            # (the matrix gradients are summed with atomics in both runs: Adam's g / sqrt(v) amplifies the
            # last-bit differences of near-zero gradient entries)
            np.testing.assert_allclose(pb.detach().cpu().numpy(), pa.detach().cpu().numpy(), rtol=2e-3, atol=2e-5, err_msg=na)
    # the fused no-grad path reads the flushed tables: same loss from both models
    f = case.formula("2-inter", cls=gqe.Formula)
    random.seed(7)
    la = dense_model.margin_loss(f, case.queries("2-inter", cls=gqe.Query)).item() if False else None
    with torch.no_grad():
        random.seed(7)
        la = dense_model.margin_loss(f, case.queries("2-inter", cls=gqe.Query)).item()
        random.seed(7)
        lb = sparse_model.margin_loss(f, case.queries("2-inter", cls=gqe.Query)).item()
    assert abs(la - lb) < 1e-4


def test_sparse_adam_equals_dense_adam_on_equal_gradients():
    """The optimiser alone: identical gradients into torch.optim.Adam (dense, zero rows included) and
    into gqe_adam_rows (touched rows only) over 40 steps with rows coming and going."""
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    n, d, lr = 300, 64, 0.01
    table = torch.randn(n, d, device=dev, generator=g) * 0.1
    ref = table.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=lr)
    m, v = torch.zeros(n, d, device=dev), torch.zeros(n, d, device=dev)
    last = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx = gqe.Context(0, torch.cuda.current_stream().cuda_stream)
    step = 0
    for it in range(40):
        if it % 7 == 3:
            continue                                   # a step in which this table gets no gradient at all
        k = int(torch.randint(1, 40, (1,), device=dev, generator=g))
        rows = torch.randperm(n, device=dev, generator=g)[:k].sort().values
        vals = torch.randn(k, d, device=dev, generator=g) * (10.0 ** float(torch.randint(-4, 1, (1,), device=dev, generator=g)))
        step += 1
        # before the "forward": the rows about to be read are caught up (what the hook does)
        ctx.adam_rows_device(table.data_ptr(), m.data_ptr(), v.data_ptr(), last.data_ptr(), n, d, k, rows.data_ptr(), None,
                             step - 1, lr, 0.9, 0.999, 1e-8)
        np.testing.assert_allclose(table[rows].cpu().numpy(), ref.detach()[rows].cpu().numpy(), rtol=1e-5, atol=1e-7)
        ctx.adam_rows_device(table.data_ptr(), m.data_ptr(), v.data_ptr(), last.data_ptr(), n, d, k, rows.data_ptr(),
                             vals.data_ptr(), step, lr, 0.9, 0.999, 1e-8)
        dense = torch.zeros(n, d, device=dev)
        dense[rows] = vals
        ref.grad = dense
        opt.step()
    ctx.adam_rows_device(table.data_ptr(), m.data_ptr(), v.data_ptr(), last.data_ptr(), n, d, n, None, None, step, lr, 0.9,
                         0.999, 1e-8)                  # flush
    torch.cuda.synchronize()
    np.testing.assert_allclose(table.cpu().numpy(), ref.detach().cpu().numpy(), rtol=1e-5, atol=2e-7)
    assert int(last.max()) == step


def test_catch_up_alone_equals_zero_gradient_adam_steps():
    """Rows that get no gradient for k steps still move under dense Adam; the catch-up replays it."""
    dev = "cuda"
    torch.manual_seed(0)
    n, d = 64, 128
    table = torch.randn(n, d, device=dev)
    ref = table.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.05)
    g0 = torch.randn(n, d, device=dev)
    m = torch.zeros(n, d, device=dev)
    v = torch.zeros(n, d, device=dev)
    last = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx = gqe.Context(0, torch.cuda.current_stream().cuda_stream)
    rows = torch.arange(0, n, 2, device=dev)                       # even rows get a gradient at step 1
    ctx.adam_rows_device(table.data_ptr(), m.data_ptr(), v.data_ptr(), last.data_ptr(), n, d, rows.numel(), rows.data_ptr(),
                         g0[rows].contiguous().data_ptr(), 1, 0.05, 0.9, 0.999, 1e-8)
    gd = torch.zeros(n, d, device=dev)
    gd[rows] = g0[rows]
    ref.grad = gd.clone()
    opt.step()
    for k in range(2, 40):                                          # 38 steps without any gradient
        ref.grad = torch.zeros(n, d, device=dev)
        opt.step()
    dup = torch.cat([rows, rows[:5]])                               # duplicates are claimed once
    ctx.adam_rows_device(table.data_ptr(), m.data_ptr(), v.data_ptr(), last.data_ptr(), n, d, dup.numel(), dup.data_ptr(),
                         None, 39, 0.05, 0.9, 0.999, 1e-8)
    torch.cuda.synchronize()
    np.testing.assert_allclose(table.cpu().numpy(), ref.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert last[rows].eq(39).all() and last[1::2].eq(0).all()
    st = opt.state[ref]
    np.testing.assert_allclose(m.cpu().numpy(), st["exp_avg"].cpu().numpy(), rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(v.cpu().numpy(), st["exp_avg_sq"].cpu().numpy(), rtol=1e-4, atol=1e-12)


@pytest.mark.parametrize("decoder,inter,d", [("bilinear", "mean", 128), ("transe", "min-simple", 64),
                                             ("bilinear-diag", "min", 32), ("bilinear", "min-simple", 32)])
def test_native_train_step_matches_dense_adam(decoder, inter, d):
    """gqe_train_step_nodes_host (NativeAdam.step: forward + backward + Adam in one native call)
    against the reference's loop body run through autograd + torch.optim.Adam: the same loss at
    every step, the same tables and operator parameters after 14 steps over all 7 structures."""
    case = make_case(seed=4, d=d, decoder=decoder, inter=inter, n_queries=160, n_neg=3, nodes_per_mode=400)
    dense_model = build_package_model(case)
    native_model = copy.deepcopy(dense_model)
    dense_opt = torch.optim.Adam(dense_model.parameters(), lr=0.01)
    native_opt = gqe.NativeAdam(native_model, lr=0.01)
    order = ["1-chain", "3-inter", "2-chain", "3-inter_chain", "1-chain", "2-inter", "3-chain_inter", "3-chain",
             "1-chain", "3-inter", "2-inter", "1-chain", "3-inter_chain", "2-chain"]
    for it, s in enumerate(order):
        f = case.formula(s, cls=gqe.Formula)
        qs = case.queries(s, cls=gqe.Query)
        lo = (it * 37) % 100
        batch = qs[lo:lo + 48]
        hard = "inter" in s and it % 2 == 1
        random.seed(1000 + it)
        dense_opt.zero_grad()
        loss = dense_model.margin_loss(f, batch, hard_negatives=hard)
        loss.backward()
        dense_opt.step()
        random.seed(1000 + it)
        got = native_opt.step(f, batch, hard_negatives=hard)
        assert abs(loss.item() - got) <= 2e-5 * max(1.0, abs(got)), (it, s, loss.item(), got)
    ctx = native_model.context()
    assert sum(ctx.train_steps(i) for i in range(len(case.kg.modes))) > 0
    native_opt.flush()
    torch.cuda.synchronize()
    for m in case.kg.modes:
        a, b = dense_model.enc.table(m).detach(), native_model.enc.table(m).detach()
        assert torch.isfinite(b).all()
        close = torch.isclose(b, a, rtol=1e-3, atol=1e-5)
        assert close.float().mean().item() > 0.99, m          # (see test_sparse_row_adam_matches_dense_adam)
        assert (a - b).abs().max().item() < 0.01 * len(order), m
        assert (a - case.tables[m].to(a.device)).abs().max().item() > 1e-3, "the tables did not train"
    for (na, pa), (nb, pb) in zip(dense_model.named_parameters(), native_model.named_parameters()):
        if "feat-" not in na:
            # (matrix gradients are summed with atomics in both runs; Adam's g / sqrt(v) amplifies the last-bit
            # differences of near-zero gradient entries: a few entries per thousand may sit within one step)
            a, b = pa.detach(), pb.detach()
            assert torch.isclose(b, a, rtol=2e-3, atol=2e-5).float().mean().item() > 0.998, na
            assert (a - b).abs().max().item() < 0.01 * len(order), na
    # the fused scoring path sees the updated operators (its packed-weight cache was dropped)
    f = case.formula("2-inter", cls=gqe.Formula)
    with torch.no_grad():
        random.seed(7)
        la = dense_model.margin_loss(f, case.queries("2-inter", cls=gqe.Query)).item()
        random.seed(7)
        lb = native_model.margin_loss(f, case.queries("2-inter", cls=gqe.Query)).item()
    assert abs(la - lb) < 1e-4
    # a node id that is not in the graph: reported like every other call
    bad = case.queries("1-chain", cls=gqe.Query)[:8]
    bad[3].anchor_nodes = (10 ** 8,)
    with pytest.raises((KeyError, IndexError)):
        native_opt.step(case.formula("1-chain", cls=gqe.Formula), bad)
    native_opt.reset()


def test_native_multitask_iteration_matches_weighted_loss_backward():
    """The reference's multi-task iteration (train_helpers.py:63-79): the edge batch plus path_weight /
    inter_weight times other query types (intersections also with hard negatives), ONE backward() and
    ONE optimizer.step().  NativeAdam.backward(weight=...) x k + apply() against autograd + optim.Adam."""
    case = make_case(seed=6, d=64, decoder="bilinear", inter="mean", n_queries=120, n_neg=3, nodes_per_mode=300)
    dense_model = build_package_model(case)
    native_model = copy.deepcopy(dense_model)
    dense_opt = torch.optim.Adam(dense_model.parameters(), lr=0.01)
    native_opt = gqe.NativeAdam(native_model, lr=0.01)
    plan = [("1-chain", False, 1.0), ("2-chain", False, 0.01), ("3-chain", False, 0.01), ("2-inter", False, 0.005),
            ("2-inter", True, 0.005), ("3-inter_chain", False, 0.005), ("3-inter_chain", True, 0.005)]
    for it in range(5):
        random.seed(50 + it)
        dense_opt.zero_grad()
        total = 0.0
        parts = []
        for s, hard, w in plan:
            f, qs = case.formula(s, cls=gqe.Formula), case.queries(s, cls=gqe.Query)
            l = dense_model.margin_loss(f, qs[it * 10:it * 10 + 40], hard_negatives=hard)
            parts.append(float(l))
            total = total + w * l
        total.backward()
        dense_opt.step()
        random.seed(50 + it)
        got = []
        for s, hard, w in plan:
            f, qs = case.formula(s, cls=gqe.Formula), case.queries(s, cls=gqe.Query)
            got.append(native_opt.backward(f, qs[it * 10:it * 10 + 40], hard_negatives=hard, weight=w))
        native_opt.apply()
        np.testing.assert_allclose(got, parts, rtol=2e-5, atol=2e-6)
    f, qs = case.formula("1-chain", cls=gqe.Formula), case.queries("1-chain", cls=gqe.Query)
    random.seed(99)
    native_opt.backward(f, qs[:8])
    with pytest.raises(gqe.GqeError):          # a pending backward blocks the one-call step
        native_opt.step(f, qs[:8])
    native_opt.apply()
    random.seed(99)                            # (the dense model takes the same extra step)
    dense_opt.zero_grad()
    dense_model.margin_loss(f, qs[:8]).backward()
    dense_opt.step()
    native_opt.flush()
    torch.cuda.synchronize()
    for m in case.kg.modes:
        a, b = dense_model.enc.table(m).detach(), native_model.enc.table(m).detach()
        assert torch.isclose(b, a, rtol=1e-3, atol=1e-5).float().mean().item() > 0.99, m
    for (na, pa), (nb, pb) in zip(dense_model.named_parameters(), native_model.named_parameters()):
        if "feat-" not in na:
            assert torch.isclose(pb.detach(), pa.detach(), rtol=2e-3, atol=2e-5).float().mean().item() > 0.995, na
