"""``QueryStore``: the reference's query files as flat integer arrays.

The reference keeps a training / test set as nested dicts of Python ``Query``
objects (netquery/data_utils.py:10-35) and walks them per query on every batch:
``run_batch`` slices a list (train_helpers.py:95-107), ``margin_loss`` calls
``random.choice`` once per query (model.py:116-120) and ``forward`` builds one
Python list per operand (model.py:75-92).  At GPU scoring rates that host work is
the whole step.  Here a file is parsed ONCE into one block per formula:

    anchors   int32 [A, Q]      anchor node ids, slot-major
    targets   int32 [Q]         positive target node ids
    negs      int32 [sum]       + neg_ptr  int64 [Q+1]   (CSR: stored negatives of query i)
    hards     int32 [sum]       + hard_ptr int64 [Q+1]   (CSR: stored hard negatives)

so that a batch is a set of array VIEWS, the negative draw is one vectorised
``rng.integers`` over the CSR lengths, and ``QueryEncoderDecoder.margin_loss`` /
``forward`` take the slice without touching a Python object per query.  The same
semantics as the reference where it matters for results: formula sampling uses the
same ``np.random.multinomial`` draw and the same wrapping window
(train_helpers.py:96-105); negatives are uniform over the same stored lists
(``reference_stream=True`` additionally consumes the global ``random`` stream exactly
like ``random.choice`` per query, for bit-for-bit comparisons with the reference).

``QueryStore.to_device(device)`` uploads every block once; slices of the resulting
``DeviceQueryStore`` are descriptors of device memory, and ``margin_loss`` on them is one
native call (``gqe_margin_loss_store_device``): a small kernel gathers the slices and draws the
negatives on the GPU, the fused scoring kernel follows -- no index array crosses PCIe per step.
"""
import pickle
import random

import numpy as np

from .query import Formula, QueryBatch, parse_query_graph


def batch_window(iter_count, batch_size, n):
    """[start, stop) of batch ``iter_count`` over ``n`` queries: contiguous, wrapping to a
    short batch at the end of the list like train_helpers.py:101-104."""
    start = (iter_count * batch_size) % n
    stop = ((iter_count + 1) * batch_size) % n
    if stop <= start or stop > n:
        stop = n
    return start, stop


def _csr(lists, n):
    ptr = np.zeros(n + 1, dtype=np.int64)
    for i, lst in enumerate(lists):
        ptr[i + 1] = ptr[i] + (0 if lst is None else len(lst))
    vals = np.empty(int(ptr[-1]), dtype=np.int32)
    for i, lst in enumerate(lists):
        if lst:
            vals[ptr[i]:ptr[i + 1]] = lst
    return ptr, vals


class FormulaBlock(object):
    """Every stored query of one formula, column-wise."""

    __slots__ = ("formula", "anchors", "targets", "neg_ptr", "negs", "hard_ptr", "hards")

    def __init__(self, formula, anchors, targets, neg_ptr, negs, hard_ptr, hards):
        self.formula, self.anchors, self.targets = formula, anchors, targets
        self.neg_ptr, self.negs, self.hard_ptr, self.hards = neg_ptr, negs, hard_ptr, hards

    def __len__(self):
        return self.targets.shape[0]

    def window(self, start, stop):
        return StoreSlice(self, int(start), int(stop))

    def all(self):
        return StoreSlice(self, 0, len(self))


def device_draw(seed, positions, lens):
    """The pick of ``gqe_store_batch`` (csrc/gqe_rows.cu) restated on the host: for the query at position
    ``positions[i]`` of a call with ``seed``, the index floor(u * lens[i]) into its negative list, u = the upper 32
    bits of splitmix64(seed + golden * (position + 1)) / 2^32.  int64 array; for tests and for reproducing a
    device-side draw on the host."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed & (2 ** 64 - 1)) + np.uint64(0x9E3779B97F4A7C15) * (np.asarray(positions, dtype=np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        u = z >> np.uint64(32)
        return ((u * np.asarray(lens, dtype=np.uint64)) >> np.uint64(32)).astype(np.int64)


class DeviceBlock(object):
    """A ``FormulaBlock`` uploaded to a GPU (torch int32 / int64 tensors; the host block stays
    reachable as ``.host``)."""

    __slots__ = ("formula", "host", "anchors", "targets", "neg_ptr", "negs", "hard_ptr", "hards", "n")

    def __init__(self, block, device):
        import torch
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.formula, self.host, self.n = block.formula, block, len(block)
        self.anchors, self.targets = up(block.anchors), up(block.targets)
        self.neg_ptr, self.negs = up(block.neg_ptr), up(block.negs if block.negs.size else np.zeros(1, np.int32))
        self.hard_ptr, self.hards = up(block.hard_ptr), up(block.hards if block.hards.size else np.zeros(1, np.int32))

    def __len__(self):
        return self.n

    def window(self, start, stop):
        return DeviceSlice(self, int(start), int(stop))

    def all(self):
        return DeviceSlice(self, 0, self.n)


class DeviceSlice(object):
    """Queries [start, stop) of a ``DeviceBlock``: a descriptor, nothing is copied or drawn on the
    host.  Accepted by ``QueryEncoderDecoder.margin_loss`` / ``margin_loss_mix``."""

    __slots__ = ("block", "start", "stop")

    def __init__(self, block, start, stop):
        if not 0 <= start <= stop <= block.n:
            raise IndexError("slice [%d, %d) outside a block of %d queries" % (start, stop, block.n))
        self.block, self.start, self.stop = block, start, stop

    def __len__(self):
        return self.stop - self.start

    @property
    def formula(self):
        return self.block.formula

    def host(self):
        """The same queries as a host ``StoreSlice``."""
        return StoreSlice(self.block.host, self.start, self.stop)


class StoreSlice(object):
    """Queries [start, stop) of a block -- views, nothing is copied.  Accepted wherever the
    scorer takes a list of ``Query`` objects of one formula."""

    __slots__ = ("block", "start", "stop")

    def __init__(self, block, start, stop):
        self.block, self.start, self.stop = block, start, stop

    def __len__(self):
        return self.stop - self.start

    @property
    def formula(self):
        return self.block.formula

    @property
    def anchors(self):
        return self.block.anchors[:, self.start:self.stop]

    @property
    def targets(self):
        return self.block.targets[self.start:self.stop]

    def _pool(self, hard):
        b = self.block
        ptr, vals = (b.hard_ptr, b.hards) if hard else (b.neg_ptr, b.negs)
        return ptr[self.start:self.stop + 1], vals

    def negative_lists(self, hard=False):
        """(offsets int64 [n+1] rebased to 0, values view): ALL stored negatives of the slice, the
        ragged layout ``eval_perc_queries`` scores (utils.py:70-91)."""
        ptr, vals = self._pool(hard)
        return ptr - ptr[0], vals[ptr[0]:ptr[-1]]

    def draw_negatives(self, hard=False, full_list=None, rng=None, reference_stream=False):
        """One negative per query -> int32 [n].  model.py:116-120: a hard negative, or (1-chain)
        any node of the target mode (``full_list``), or one of the query's stored negatives.
        ``rng``: a ``numpy.random.Generator`` (default: a process-wide one).
        ``reference_stream``: draw with the global ``random`` module, one call per query in
        query order, consuming exactly what the reference's ``random.choice`` calls consume."""
        n = len(self)
        if full_list is not None and not hard:
            pool = np.asarray(full_list)
            if reference_stream:
                pick = np.fromiter((random.randrange(len(pool)) for _ in range(n)), dtype=np.int64, count=n)
            else:
                pick = (rng or _default_rng()).integers(0, len(pool), size=n)
            return pool[pick].astype(np.int32)
        ptr, vals = self._pool(hard)
        lens = ptr[1:] - ptr[:-1]
        if n and int(lens.min()) <= 0:
            raise IndexError("a query of the batch has no %snegative samples" % ("hard " if hard else ""))
        if reference_stream:
            pick = np.fromiter((random.randrange(int(k)) for k in lens), dtype=np.int64, count=n)
        else:
            # uniform over [0, len_i) per query from ONE block of 32-bit words: (r * len) >> 32 (the
            # bias, len / 2^32, is far below anything a training run resolves); Generator.integers with
            # an array of bounds costs 12 ns per element, 0.8 ms for a 65 536-query batch
            r = (rng or _default_rng()).integers(0, 1 << 32, size=n, dtype=np.uint32)
            pick = np.multiply(r, lens.astype(np.uint32), dtype=np.uint64)
            np.right_shift(pick, 32, out=pick)
            pick = pick.view(np.int64)
        return vals[ptr[:-1] + pick]

    def margin_batch(self, negatives):
        """-> QueryBatch of (positive, negative) pairs, int32 node ids."""
        pairs = np.empty((len(self), 2), dtype=np.int32)
        pairs[:, 0] = self.targets
        pairs[:, 1] = negatives
        return QueryBatch(self.formula, self.anchors, pairs.reshape(-1))


_RNG = None


def _default_rng():
    global _RNG
    if _RNG is None:
        _RNG = np.random.default_rng()
    return _RNG


class QueryStore(object):
    """{Formula: FormulaBlock} plus the by-type index the training loop walks
    (``train_queries[query_type]``, train_helpers.py:51,64-72)."""

    def __init__(self, blocks):
        self.blocks = dict(blocks)
        self.by_type = {}
        for f in self.blocks:
            self.by_type.setdefault(f.query_type, []).append(f)

    def __len__(self):
        return sum(len(b) for b in self.blocks.values())

    def formulas(self, query_type=None):
        return list(self.blocks) if query_type is None else list(self.by_type.get(query_type, ()))

    def __getitem__(self, formula):
        return self.blocks[formula]

    def to_device(self, device):
        """Upload every block once -> ``DeviceQueryStore`` (same sampling interface; its slices
        never touch host memory again)."""
        return DeviceQueryStore(self, device)

    # ---- construction ---------------------------------------------------------
    @classmethod
    def from_records(cls, records, select=None):
        """``records``: the reference's serialised queries ``(query_graph, neg_samples,
        hard_neg_samples)`` (graph.py:93-96), in file order.  ``select(record)`` filters."""
        groups = {}
        for rec in records:
            if select is not None and not select(rec):
                continue
            qt, rels, target, anchors = parse_query_graph(rec[0])
            groups.setdefault((qt, rels), []).append((target, anchors, rec[1], rec[2]))
        blocks = {}
        for (qt, rels), rows in groups.items():
            f = Formula(qt, rels)
            n = len(rows)
            anchors = np.array([r[1] for r in rows], dtype=np.int32).reshape(n, len(f.anchor_modes)).T.copy()
            targets = np.fromiter((r[0] for r in rows), dtype=np.int32, count=n)
            neg_ptr, negs = _csr([r[2] for r in rows], n)
            hard_ptr, hards = _csr([r[3] for r in rows], n)
            blocks[f] = FormulaBlock(f, anchors, targets, neg_ptr, negs, hard_ptr, hards)
        return cls(blocks)

    @classmethod
    def from_file(cls, path, select=None):
        """A query pickle written by the reference (Python 2: ``encoding="latin1"``)."""
        with open(path, "rb") as fh:
            return cls.from_records(pickle.load(fh, encoding="latin1"), select)

    @classmethod
    def test_split(cls, path):
        """data_utils.py:27-35: {"full_neg": store, "one_neg": store} by whether more than one
        negative was stored with the query."""
        with open(path, "rb") as fh:
            records = pickle.load(fh, encoding="latin1")
        return {"full_neg": cls.from_records(records, lambda r: len(r[1]) > 1),
                "one_neg": cls.from_records(records, lambda r: len(r[1]) <= 1)}

    # ---- batch sampling (train_helpers.py:95-107) ---------------------------------
    def sample_batch(self, query_type, iter_count, batch_size):
        """ONE formula of ``query_type`` drawn ~ multinomial(#queries per formula) from numpy's
        global RNG (the same ``np.random.multinomial(1, p)`` call, so the same stream as the
        reference), then the wrapping window of its block -> StoreSlice."""
        formulas = self.by_type[query_type]
        sizes = np.array([len(self.blocks[f]) for f in formulas], dtype=np.float64)
        which = int(np.argmax(np.random.multinomial(1, sizes / sizes.sum())))
        block = self.blocks[formulas[which]]
        return block.window(*batch_window(iter_count, batch_size, len(block)))


class DeviceQueryStore(QueryStore):
    """A ``QueryStore`` whose blocks live in GPU memory (``QueryStore.to_device``).  ``sample_batch``
    draws the formula exactly like the host store (same ``np.random.multinomial`` call) and returns
    a ``DeviceSlice``."""

    def __init__(self, store, device):
        self.host = store
        self.device = device
        self.blocks = {f: DeviceBlock(b, device) for f, b in store.blocks.items()}
        self.by_type = {t: list(fs) for t, fs in store.by_type.items()}
