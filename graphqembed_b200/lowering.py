"""Host lowering: formulas -> gqe_plan, node ids -> table rows.

This is the vectorised replacement for the per-node Python work on the
reference's path: the ``node_maps[mode][n]`` dict lookups inside the
``features`` closure (reference netquery/bio/data_utils.py:20-21) and the
per-operand list comprehensions of ``QueryEncoderDecoder.forward``
(model.py:75,80,83,91).  The *values* produced are identical (same rows, same
relation keys in the same order); tests compare them with a trace of the
reference.
"""
import numpy as np

from . import _lib
from .query import CHAIN_TYPES, FLAT_INTER_TYPES, reverse_relation


class RowLookup(object):
    """node id -> embedding-table row, per mode:  row = node_maps[mode][n] + 1.

    Callable like the reference's ``features`` closure argument of
    ``DirectEncoder`` but returning row indices instead of embeddings, so that
    the gather can be fused into the scoring kernel.  ``node_maps`` is either
    the reference's ``{mode: {node: position}}`` dicts, ``{mode: id_array}``
    (position i holds the node id, as stored in graph_data.pkl), or None for
    identity (node id == position, reference utils.py:18-20).

    Each mode is held as a dense table ``lut[node - base] = row`` (-1: not a node
    of the mode): an O(1) lookup on the host (``rows``) and, uploaded once per
    device (``device_maps``), the table the CUDA kernels index themselves when
    they are handed node ids (``gqe_bind_node_maps``).  A mode whose ids are too
    sparse for a dense table (range > 16 x nodes + 2^20) falls back to a sorted
    search on the host only.
    """

    DENSE_SLACK = 16
    DENSE_FLOOR = 1 << 20

    def __init__(self, node_maps=None):
        self._lut = {}       # mode -> (base, int32 lut)
        self._sorted = {}    # mode -> (keys, vals)  (sparse ids only)
        self._device = {}    # device -> (modes, [lut tensors])
        self.identity = node_maps is None
        if node_maps is None:
            return
        for mode, m in node_maps.items():
            if isinstance(m, dict):
                keys = np.fromiter(m.keys(), dtype=np.int64, count=len(m))
                vals = np.fromiter(m.values(), dtype=np.int64, count=len(m))
            else:
                keys = np.asarray(m, dtype=np.int64)
                vals = np.arange(len(keys), dtype=np.int64)
            if len(keys) == 0:
                self._lut[mode] = (0, np.full(1, -1, dtype=np.int32))
                continue
            lo, hi = int(keys.min()), int(keys.max())
            if hi - lo + 1 <= self.DENSE_SLACK * len(keys) + self.DENSE_FLOOR and vals.max() + 1 < 2 ** 31:
                lut = np.full(hi - lo + 1, -1, dtype=np.int32)
                lut[keys - lo] = (vals + 1).astype(np.int32)
                self._lut[mode] = (lo, lut)
            else:
                order = np.argsort(keys, kind="stable")
                self._sorted[mode] = (keys[order], vals[order])

    @property
    def modes(self):
        return list(self._lut) + list(self._sorted)

    def rows(self, nodes, mode):
        """int32 table rows for ``nodes`` (any int sequence / array) of ``mode``.
        KeyError on an unknown mode or node, like the reference's dict lookups."""
        nodes = np.asarray(nodes, dtype=np.int64)
        if self.identity:
            return (nodes + 1).astype(np.int32)
        if mode in self._lut:
            base, lut = self._lut[mode]
            k = nodes - base
            inside = (k >= 0) & (k < len(lut))
            r = lut[np.where(inside, k, 0)]
            bad = ~inside | (r < 0)
            if bad.any():
                raise KeyError(int(nodes.reshape(-1)[np.flatnonzero(bad.reshape(-1))[0]]))
            return r
        keys, vals = self._sorted[mode]          # KeyError on an unknown mode, like the reference
        pos = np.searchsorted(keys, nodes)
        pos_c = np.minimum(pos, len(keys) - 1)
        bad = keys[pos_c] != nodes
        if bad.any():
            raise KeyError(int(nodes.reshape(-1)[np.flatnonzero(bad.reshape(-1))[0]]))
        return (vals[pos_c] + 1).astype(np.int32)

    __call__ = rows

    def device_maps(self, modes, table_rows, device):
        """-> (lut device pointers, bases, lens, keep-alive tensors) for ``gqe_bind_node_maps``
        in the order of ``modes``, or None when some mode has no dense table (sparse ids: the
        host lookup is used instead).  Identity maps are affine: row = node + 1."""
        if self.identity:
            return [0] * len(modes), [-1] * len(modes), [int(r) for r in table_rows], []
        if any(m not in self._lut for m in modes):
            return None
        import torch
        key = (str(device), tuple(modes))
        held = self._device.get(key)
        if held is None:
            held = [torch.from_numpy(self._lut[m][1]).to(device) for m in modes]
            self._device[key] = held
        return ([t.data_ptr() for t in held], [self._lut[m][0] for m in modes], [self._lut[m][1].size for m in modes],
                held)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_device"] = {}     # device copies are rebuilt on demand
        return state


def relation_order(formula):
    """Relation keys in the order the reference touches its parameter dict.

    chains: r1..rn on the target (decoders.py:143-145); intersections: the
    REVERSED relations on the anchors (model.py:81-92), the nested branch
    walking ``rels[1][::-1]`` (model.py:84-86); 3-chain_inter: the two reversed
    branch relations, then reverse(r1) after the intersection (model.py:102-107).
    """
    qt, rels = formula.query_type, formula.rels
    if qt in CHAIN_TYPES:
        return list(rels)
    if qt in FLAT_INTER_TYPES:
        return [reverse_relation(r) for r in rels]
    if qt == "3-inter_chain":
        return [reverse_relation(rels[0]), reverse_relation(rels[1][1]), reverse_relation(rels[1][0])]
    if qt == "3-chain_inter":
        return [reverse_relation(rels[1][0]), reverse_relation(rels[1][1]), reverse_relation(rels[0])]
    raise ValueError("unsupported query type %r" % (qt,))


def inter_mode_of(formula):
    """Mode whose pre/post matrices the intersection uses (model.py:95,106)."""
    if formula.query_type in CHAIN_TYPES:
        return None
    if formula.query_type == "3-chain_inter":
        return formula.rels[0][-1]
    return formula.target_mode


def lower_formula(formula, mode_ids, rel_ids):
    """Formula -> ctypes ``Plan`` (gqe_plan).  ``mode_ids`` / ``rel_ids`` map
    mode names / canonical relation triples to the ids the context was bound
    with; an unknown key raises KeyError like the reference's dict lookups."""
    plan = _lib.Plan()
    plan.structure = _lib.STRUCTURE_ID[formula.query_type]
    plan.target_mode = mode_ids[formula.target_mode]
    for k in range(_lib.GQE_MAX_ANCHORS):
        plan.anchor_mode[k] = mode_ids[formula.anchor_modes[k]] if k < len(formula.anchor_modes) else -1
    im = inter_mode_of(formula)
    plan.inter_mode = -1 if im is None else mode_ids[im]
    order = relation_order(formula)
    for k in range(_lib.GQE_MAX_RELS):
        plan.rel[k] = rel_ids[order[k]] if k < len(order) else -1
    return plan
