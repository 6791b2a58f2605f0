"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference modules in this container.

Imports ``netquery.model`` / ``decoders`` / ``encoders`` straight from the
read-only reference tree (default ``/root/reference``) under Python 3 /
torch 2.x, without editing or copying any reference file:

* ``netquery.graph`` contains Python-2 ``print`` statements further down, so a
  stub module is registered that executes only the file's head (everything
  before ``class Graph``: ``_reverse_relation``, ``Formula``, ``Query``) read
  from the reference tree at run time;
* ``DirectEncoder.__init__`` calls ``dict.iteritems`` -> a dict subclass;
* ``agg_func=torch.min`` returns a named tuple the reference's
  ``type(x) == tuple`` test misses on modern torch -> a values-only lambda.

The reference tree does not exist on the GPU box: nothing imported by the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may use this module.  It is used
by ``oracle/make_golden.py`` and by the CPU test that pins the oracle.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GQE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "netquery", "model.py"))


_mods = None


def load():
    """-> (graph_head_module, netquery.model, netquery.decoders, netquery.encoders)"""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with open(os.path.join(REFERENCE_ROOT, "netquery", "graph.py")) as fh:
        src = fh.read()
    head = src[:src.index("class Graph()")]
    import netquery
    g = types.ModuleType("netquery.graph")
    exec(compile(head, "netquery/graph.py[head]", "exec"), g.__dict__)
    sys.modules["netquery.graph"] = g
    netquery.graph = g
    import netquery.decoders as D
    import netquery.encoders as E
    import netquery.model as M
    _mods = (g, M, D, E)
    return _mods


class IterDict(dict):
    def iteritems(self):
        return self.items()


class TracingDict(dict):
    """Parameter dict that logs the order in which keys are read."""

    def __init__(self, base, log):
        dict.__init__(self, base)
        self._log = log

    def __getitem__(self, key):
        self._log.append(("rel", key))
        return dict.__getitem__(self, key)


class GraphLike(object):
    def __init__(self, full_lists):
        self.full_lists = full_lists


def build_reference_model(tables, node_maps, relations, rel_params, decoder, inter, pre, post, full_lists=None,
                          trace=None):
    """The reference's QueryEncoderDecoder carrying the given parameter values.

    tables / rel_params / pre / post are {key: tensor}; ``relations`` is the
    reference's ``{mode: [(to_mode, name), ...]}``.  When ``trace`` is a list it
    receives ("rows", mode, [...]) for every embedding lookup and ("rel", key)
    for every relation-parameter read, in call order.
    """
    g, M, D, E = load()
    modes = list(tables.keys())
    d = next(iter(tables.values())).size(1)
    dims = {m: d for m in modes}
    feature_modules = IterDict()
    for m in modes:
        emb = torch.nn.Embedding(tables[m].size(0), d)
        emb.weight.data.copy_(tables[m])
        feature_modules[m] = emb

    def features(nodes, mode):   # shape of reference netquery/bio/data_utils.py:20-21
        idx = torch.LongTensor([node_maps[mode][n] for n in nodes]) + 1
        if trace is not None:
            trace.append(("rows", mode, idx.tolist()))
        return feature_modules[mode](idx)

    enc = E.DirectEncoder(features, feature_modules)
    if decoder == "bilinear":
        dec = D.BilinearMetapathDecoder(relations, dims)
        store = dec.mats
    elif decoder == "transe":
        dec = D.TransEMetapathDecoder(relations, dims)
        store = dec.vecs
    elif decoder == "bilinear-diag":
        dec = D.BilinearDiagMetapathDecoder(relations, dims)
        store = dec.vecs
    else:
        raise ValueError(decoder)
    for rel, p in rel_params.items():
        store[rel].data.copy_(p)
    if trace is not None:
        logged = TracingDict(store, trace)
        if decoder == "bilinear":
            dec.mats = logged
        else:
            dec.vecs = logged
    amin = lambda x, dim: torch.min(x, dim=dim)[0]
    agg = torch.mean if inter.startswith("mean") else amin
    if inter.endswith("-simple"):
        idec = D.SimpleSetIntersection(agg_func=agg)
    else:
        idec = D.SetIntersection(dims, dims, agg_func=agg)
        for m in modes:
            idec.pre_mats[m].data.copy_(pre[m])
            idec.post_mats[m].data.copy_(post[m])
    model = M.QueryEncoderDecoder(GraphLike(full_lists), enc, dec, idec)
    return model, g


def reference_eval_functions():
    """``eval_auc_queries`` / ``eval_perc_queries`` / ``_get_perc_scores`` of the reference,
    executed from its own source: netquery/utils.py cannot be imported under py3 (``cPickle``,
    utils.py:9), so lines 26-91 are exec'd verbatim in a namespace that supplies the names the
    module header would have imported (and py2's ``xrange``)."""
    import random as _random

    import numpy as _np
    from scipy import stats as _stats
    from sklearn.metrics import roc_auc_score as _auc
    src = open(os.path.join(REFERENCE_ROOT, "netquery", "utils.py")).read().split("\n")
    body = "\n".join(src[25:91])
    ns = {"np": _np, "stats": _stats, "roc_auc_score": _auc, "random": _random, "xrange": range}
    exec(compile(body, "netquery/utils.py[26:91]", "exec"), ns)
    return ns["eval_auc_queries"], ns["eval_perc_queries"]


def reference_run_batch():
    """``run_batch`` of the reference (netquery/train_helpers.py:95-107) executed from its own
    source.  The module cannot be imported under py3 (implicit relative import, ``xrange``), and
    line 100 indexes ``dict.keys()`` -- callers pass a dict whose ``keys()`` returns a list."""
    import numpy as _np
    src = open(os.path.join(REFERENCE_ROOT, "netquery", "train_helpers.py")).read().split("\n")
    ns = {"np": _np}
    exec(compile("\n".join(src[94:107]), "netquery/train_helpers.py[95:107]", "exec"), ns)
    return ns["run_batch"]
