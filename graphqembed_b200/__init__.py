"""graphqembed_b200 -- B200-native conjunctive-query embedding scorer.

A from-scratch sm_100a implementation of ONE path of williamleif/graphqembed
(``netquery``): the batched forward scoring of conjunctive graph queries and
its margin loss, behind the reference's own operator surface.  See DESIGN.md
for the path and its boundary, INTEGRATION.md for how it plugs into netquery.

Importing this package does not need a GPU; creating a context (any scoring
call) does, and raises if the CUDA library or an sm_100 device is missing.
"""
from . import _lib
from ._lib import Context, GqeError, GqeIndexError, Plan, Segment, build, load, make_segments
from .lowering import RowLookup, lower_formula, relation_order
from .query import Formula, Query, QueryBatch, reverse_relation
from .store import DeviceQueryStore, DeviceSlice, FormulaBlock, QueryStore, StoreSlice

__all__ = ["Context", "GqeError", "GqeIndexError", "QueryStore", "StoreSlice", "FormulaBlock", "DeviceQueryStore", "DeviceSlice", "SparseRowAdam", "NativeAdam", "Plan", "Segment", "build", "load", "make_segments", "RowLookup", "lower_formula",
           "relation_order", "Formula", "Query", "QueryBatch", "reverse_relation", "DirectEncoder",
           "BilinearMetapathDecoder", "TransEMetapathDecoder", "BilinearDiagMetapathDecoder", "SetIntersection",
           "SimpleSetIntersection", "QueryEncoderDecoder", "get_encoder", "get_metapath_decoder",
           "get_intersection_decoder", "eval_auc_queries", "eval_perc_queries", "load_graph", "load_queries",
           "load_queries_by_formula", "load_queries_by_type", "load_test_queries_by_formula", "run_batch", "pick_batch"]

_TORCH_SIDE = {
    "DirectEncoder": "operators", "BilinearMetapathDecoder": "operators", "TransEMetapathDecoder": "operators",
    "BilinearDiagMetapathDecoder": "operators", "SetIntersection": "operators", "SimpleSetIntersection": "operators",
    "get_encoder": "operators", "get_metapath_decoder": "operators", "get_intersection_decoder": "operators",
    "cosine_similarity_dim0": "operators", "QueryEncoderDecoder": "scorer", "SparseRowAdam": "optim", "NativeAdam": "optim",
    "eval_auc_queries": "evaluation", "eval_perc_queries": "evaluation",
    "load_graph": "data", "load_queries": "data", "load_queries_by_formula": "data", "load_queries_by_type": "data",
    "load_test_queries_by_formula": "data", "run_batch": "data", "pick_batch": "data",
}


def __getattr__(name):
    # torch-facing classes are imported lazily so that the pure C-ABI users
    # (bench, ctypes hosts) do not pay the torch import.
    mod = _TORCH_SIDE.get(name)
    if mod is None:
        raise AttributeError(name)
    import importlib
    return getattr(importlib.import_module("." + mod, __name__), name)
