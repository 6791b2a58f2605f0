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
    """

    def __init__(self, node_maps=None):
        self._keys = {}
        self._vals = {}
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
            order = np.argsort(keys, kind="stable")
            self._keys[mode] = keys[order]
            self._vals[mode] = vals[order]

    def rows(self, nodes, mode):
        """int32 table rows for ``nodes`` (any int sequence / array) of ``mode``."""
        nodes = np.asarray(nodes, dtype=np.int64)
        if self.identity:
            return (nodes + 1).astype(np.int32)
        keys = self._keys[mode]          # KeyError on an unknown mode, like the reference
        pos = np.searchsorted(keys, nodes)
        pos_c = np.minimum(pos, len(keys) - 1)
        bad = keys[pos_c] != nodes
        if bad.any():
            raise KeyError(int(nodes.reshape(-1)[np.flatnonzero(bad.reshape(-1))[0]]))
        return (self._vals[mode][pos_c] + 1).astype(np.int32)

    __call__ = rows


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
