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
import torch


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
