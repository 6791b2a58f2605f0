"""``SparseRowAdam``: the reference's ``optim.Adam`` step at the cost of the rows a batch touched.

The reference trains with dense ``torch.optim.Adam`` over ``nn.Embedding`` tables
(netquery/bio/train.py:59-62): each step builds an ``[N_mode + 2, d]`` gradient per
touched mode and rewrites every row of those tables and of both moment buffers, although
a 512-query batch gathers ~1500 rows.  This optimiser keeps the SAME trajectory --

* rows that received a gradient take the ordinary Adam step;
* rows that did not are not left behind as in ``SparseAdam``: dense Adam keeps moving a
  row after its last gradient (the moments decay step by step), and ``gqe_adam_rows``
  replays exactly those zero-gradient steps the next time the row is needed -- before the
  forward pass reads it (``attach`` hooks the model's gathers) or at ``flush()``;
* a table without any gradient in a step is skipped and its step counter does not
  advance, like ``torch.optim.Adam`` with ``zero_grad(set_to_none=True)``;

-- while touching only those rows (csrc/gqe_opt.cu).  The small dense parameters (relation
matrices / vectors, DeepSets pre / post) go through ``torch.optim.Adam`` unchanged.

    opt = SparseRowAdam(model, lr=0.01)          # instead of optim.Adam(model.parameters(), lr=0.01)
    loss = model.margin_loss(formula, queries); loss.backward(); opt.step(); opt.zero_grad()
    opt.flush()                                   # before evaluation / torch.save: all rows up to date
"""
import numpy as np
import torch

from . import _lib
from .query import QueryBatch
from .store import DeviceSlice, StoreSlice


class NativeAdam(object):
    """The whole loop body of the reference's training (netquery/train_helpers.py:76-79) as ONE
    native call per batch:

        opt = NativeAdam(model, lr=0.01)                 # instead of optim.Adam(model.parameters(), lr=0.01)
        loss = opt.step(formula, queries)                # zero_grad + margin_loss + backward + optimizer.step()
        opt.flush()                                      # before evaluation / torch.save: all rows up to date

    ``gqe_train_step_nodes_host`` runs the exact-fp32 forward pass, the backward pass and the Adam
    update (dense torch.optim.Adam trajectory; the tables are updated row-wise with exact catch-up,
    see ``SparseRowAdam``) without returning to Python between kernels.  The model's parameters are
    updated in place; the optimiser state lives in the model's native context.  ``queries``: a list of
    ``Query`` objects, a ``StoreSlice`` or a ``DeviceSlice`` (its host arrays are used); negatives are drawn exactly as
    ``margin_loss`` draws them from a host slice.
    """

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.hyper = _lib.AdamHyper(lr, betas[0], betas[1], eps)
        self._pin = None

    def _pinned(self, n_int32):
        if self._pin is None or self._pin.numel() < n_int32:
            self._pin = torch.empty(max(n_int32 + n_int32 // 2, 4096), dtype=torch.int32).pin_memory()
        return self._pin

    def step(self, formula, queries, hard_negatives=False, margin=1):
        """-> the batch's margin loss before the update (Python float)."""
        return self._run(formula, queries, hard_negatives, margin, None)

    def backward(self, formula, queries, hard_negatives=False, weight=1.0, margin=1):
        """Forward + backward of one batch, its gradients times ``weight`` ACCUMULATED on top of earlier
        ``backward`` calls, no update; -> the batch's own (unweighted) loss.  With ``apply()`` this is the
        reference's multi-task iteration (train_helpers.py:63-79):

            loss = opt.backward(f_edge, edge_batch)                                  # run_batch(train_queries["1-chain"], ...)
            loss += path_weight * opt.backward(f_path, path_batch, weight=path_weight)
            loss += inter_weight * opt.backward(f_int, int_batch, weight=inter_weight)
            loss += inter_weight * opt.backward(f_int, int_batch, hard_negatives=True, weight=inter_weight)
            opt.apply()                                                              # loss.backward(); optimizer.step()
        """
        return self._run(formula, queries, hard_negatives, margin, float(weight))

    def apply(self):
        """One Adam step on everything that received a gradient since the last ``apply`` / ``step``."""
        self.model.context().train_apply(self.hyper)

    def _run(self, formula, queries, hard_negatives, margin, weight):
        m = self.model
        ctx = m.context()
        if isinstance(queries, DeviceSlice):       # the training call takes host index arrays: the same slice of the host block
            queries = queries.host()
        if isinstance(queries, StoreSlice):
            if "inter" not in formula.query_type and hard_negatives:
                raise Exception("Hard negative examples can only be used with intersection queries")
            full = m._full_array(formula.target_mode) if formula.query_type == "1-chain" else None
            neg = queries.draw_negatives(hard_negatives, full, m.negative_rng, m.reference_negatives)
            anchors, pos = np.asarray(queries.anchors), np.asarray(queries.targets)
        else:
            neg = m.pick_negatives(formula, queries, hard_negatives)
            n = len(queries)
            anchors = np.empty((len(formula.anchor_modes), n), dtype=np.int64)
            for k in range(anchors.shape[0]):
                anchors[k] = np.fromiter((q.anchor_nodes[k] for q in queries), dtype=np.int64, count=n)
            pos = np.fromiter((q.target_node for q in queries), dtype=np.int64, count=n)
        n = len(pos)
        pairs = np.empty((n, 2), dtype=np.int64)
        pairs[:, 0] = pos
        pairs[:, 1] = np.asarray(neg)
        batch = QueryBatch(formula, anchors, pairs.reshape(-1))
        a, t, nodes = m._indices(batch)                    # node ids (device lookup) or host-lowered rows
        na = a.shape[0]
        pin = self._pinned(na * n + 2 * n)
        view = pin.numpy()
        view[:na * n] = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
        view[na * n:na * n + 2 * n] = np.ascontiguousarray(t, dtype=np.int32).reshape(-1)
        base = pin.data_ptr()
        if weight is None:
            return ctx.train_step_host(m.plan(formula), n, base, base + 4 * na * n, margin, self.hyper, nodes=nodes)
        return ctx.train_backward_host(m.plan(formula), n, base, base + 4 * na * n, margin, weight, self.hyper, nodes=nodes)

    def flush(self):
        self.model.context().train_flush()

    def reset(self):
        self.model.context().train_reset()

    def zero_grad(self, set_to_none=True):     # (nothing to clear: gradients never leave the native step)
        pass


class SparseRowAdam(object):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        enc = model.enc
        self.tables = {m: enc.table(m) for m in enc.modes}
        table_ids = set(id(t) for t in self.tables.values())
        dense = [p for p in model.parameters() if id(p) not in table_ids and p.requires_grad]
        self.dense = torch.optim.Adam(dense, lr=lr, betas=betas, eps=eps) if dense else None
        self.state = {}
        for m, t in self.tables.items():
            self.state[m] = {"step": 0, "exp_avg": torch.zeros_like(t), "exp_avg_sq": torch.zeros_like(t),
                             "last": torch.zeros(t.size(0), dtype=torch.int32, device=t.device)}
        self.attach(model)

    # the model asks for sparse table gradients and announces the rows it is about to gather
    def attach(self, model):
        model.sparse_table_grads = True
        model.row_hook = self._before_gather

    def detach(self):
        self.flush()
        self.model.sparse_table_grads = False
        self.model.row_hook = None

    def _rows_call(self, mode, n, rows_ptr, grads_ptr, step):
        st, t = self.state[mode], self.tables[mode]
        self.model.context().adam_rows_device(t.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                              st["last"].data_ptr(), t.size(0), t.size(1), n, rows_ptr, grads_ptr, step,
                                              self.lr, self.betas[0], self.betas[1], self.eps)

    def _before_gather(self, mode, rows):
        """rows: int32 CUDA tensor of table rows the forward pass will read."""
        st = self.state[mode]
        if st["step"] > 0 and rows.numel():
            rows64 = rows.to(torch.int64)
            self._rows_call(mode, rows64.numel(), rows64.data_ptr(), None, st["step"])

    def zero_grad(self, set_to_none=True):
        for t in self.tables.values():
            t.grad = None
        if self.dense is not None:
            self.dense.zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self):
        for m, t in self.tables.items():
            g = t.grad
            if g is None:
                continue                       # untouched table: skipped, its step count stands (torch.optim.Adam)
            st = self.state[m]
            st["step"] += 1
            if g.is_sparse:
                g = g.coalesce()               # unique rows, summed gradients
                rows, vals = g.indices()[0].contiguous(), g.values().contiguous()
            else:                              # a dense gradient (model not attached): every row is "touched"
                rows = torch.arange(t.size(0), device=t.device)
                vals = g.contiguous()
            self._rows_call(m, rows.numel(), rows.data_ptr(), vals.data_ptr(), st["step"])
        if self.dense is not None:
            self.dense.step()

    @torch.no_grad()
    def flush(self):
        """Bring every row of every table up to date (before evaluation, checkpoints, or handing
        the tables to a dense optimiser)."""
        for m, t in self.tables.items():
            if self.state[m]["step"] > 0:
                self._rows_call(m, t.size(0), None, None, self.state[m]["step"])
