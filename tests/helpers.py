"""Shared test plumbing: golden fixtures, package models built from a case."""
import glob
import os

import numpy as np
import torch

from oracle.cases import Case

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GOLDEN_IDS = [os.path.basename(p)[:-4] for p in GOLDEN_FILES]
LOSS_SEED = 20260917     # oracle/make_golden.py


def load_golden(path):
    return Case.load(path)


class GraphLike(object):
    def __init__(self, kg, features):
        self.full_lists = kg.full_lists()
        self.relations = kg.relations
        self.features = features


def build_package_model(case, device="cuda"):
    """graphqembed_b200 operator stack carrying the case's parameter values."""
    import graphqembed_b200 as gqe
    kg, d = case.kg, case.d
    feature_modules = {}
    for m in kg.modes:
        emb = torch.nn.Embedding(case.tables[m].size(0), d)
        emb.weight.data.copy_(case.tables[m])
        feature_modules[m] = emb
    lookup = gqe.RowLookup(kg.node_maps())
    graph = GraphLike(kg, lookup)
    dims = {m: d for m in kg.modes}
    enc = gqe.get_encoder(0, graph, dims, feature_modules)
    dec = gqe.get_metapath_decoder(graph, dims, case.decoder)
    store = dec.mats if case.decoder == "bilinear" else dec.vecs
    for rel, p in case.rel_params.items():
        store[rel].data.copy_(p)
    idec = gqe.get_intersection_decoder(graph, dims, case.inter)
    if not case.inter.endswith("-simple"):
        for m in kg.modes:
            idec.pre_mats[m].data.copy_(case.pre[m])
            idec.post_mats[m].data.copy_(case.post[m])
    model = gqe.QueryEncoderDecoder(graph, enc, dec, idec)
    return model.to(device) if device else model


def query_batch(case, structure, targets):
    """QueryBatch of the case's queries against explicit targets ([Q] or [Q,T])."""
    import graphqembed_b200 as gqe
    b = case.batches[structure]
    return gqe.QueryBatch(case.formula(structure, cls=gqe.Formula), b["anchors"], np.asarray(targets).reshape(-1))
