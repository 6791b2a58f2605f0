"""TEST INFRASTRUCTURE ONLY -- seeded parity cases (parameters + query batches).

A *case* is everything needed to evaluate the scoring path once: a synthetic
typed graph, parameter values drawn from the reference's initialisers
(tables ~ N(0, 1/d): netquery/bio/data_utils.py:17-19; relation matrices and
pre/post xavier-uniform: decoders.py:139,282,285; relation vectors
U(+-6/sqrt(d)): decoders.py:195,224) and, per query structure, one formula
with node-id batches.  Cases round-trip through ``.npz`` files so that golden
vectors produced with the real reference (``oracle/make_golden.py``) can be
replayed on a machine where the reference tree does not exist.
"""
import json
import math

import numpy as np
import torch

from graphqembed_b200.synth import N_ANCHORS, STRUCTURES, SynthKG

from . import netquery_oracle as O

DECODERS = ("bilinear", "transe", "bilinear-diag")
INTERS = ("mean", "min", "mean-simple", "min-simple")


def _tup(x):
    """JSON lists -> the nested tuples relation structures are made of."""
    return tuple(_tup(v) for v in x) if isinstance(x, list) else x


class Case(object):
    def __init__(self, kg, d, decoder, inter, tables, rel_params, pre, post, batches):
        self.kg, self.d, self.decoder, self.inter = kg, d, decoder, inter
        self.tables, self.rel_params, self.pre, self.post = tables, rel_params, pre, post
        self.batches = batches      # {structure: dict(rels, target, anchors, negs)}

    # ---- oracle-side views --------------------------------------------------
    def oracle(self, dtype=torch.float32):
        return O.OracleScorer(self.tables, self.kg.node_maps(), self.rel_params, self.decoder, self.inter,
                              self.pre, self.post, full_lists=self.kg.full_lists(), dtype=dtype)

    def queries(self, structure, cls=O.Query, max_negs=None):
        """Query objects of ``cls`` (oracle / package / reference Query class)."""
        b = self.batches[structure]
        out = []
        for i in range(len(b["target"])):
            qg = SynthKG.query_graph(structure, b["rels"], b["target"][i], b["anchors"][:, i])
            negs = [int(x) for x in b["negs"][i]]
            if max_negs is not None:
                negs = negs[:max_negs]
            # deserialize-style construction: keep the lists exactly as given
            out.append(cls(qg, negs, negs, len(negs) + 1))
        return out

    def formula(self, structure, cls=O.Formula):
        return cls(structure, self.batches[structure]["rels"])

    # ---- (de)serialisation ----------------------------------------------------
    def save(self, path, expected=None):
        arrs = {}
        meta = {
            "modes": self.kg.modes, "sizes": [self.kg.sizes[m] for m in self.kg.modes],
            "relations": {m: [list(r) for r in v] for m, v in self.kg.relations.items()},
            "d": self.d, "decoder": self.decoder, "inter": self.inter,
            "rels": {s: self.batches[s]["rels"] for s in self.batches},
        }
        for m in self.kg.modes:
            arrs["ids/" + m] = self.kg.node_ids[m]
            arrs["table/" + m] = self.tables[m].numpy()
            if self.pre:
                arrs["pre/" + m] = self.pre[m].numpy()
                arrs["post/" + m] = self.post[m].numpy()
        for i, rel in enumerate(self.kg.rel_keys):
            arrs["rel/%d" % i] = self.rel_params[rel].numpy()
        for s, b in self.batches.items():
            for k in ("target", "anchors", "negs"):
                arrs["q/%s/%s" % (s, k)] = b[k]
        for k, v in (expected or {}).items():
            if isinstance(v, np.ndarray):
                arrs["exp/" + k] = v
            else:
                meta.setdefault("exp", {})[k] = v
        arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(path, **arrs)

    @staticmethod
    def load(path):
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())
        kg = SynthKG.__new__(SynthKG)
        kg.modes = meta["modes"]
        kg.sizes = dict(zip(kg.modes, meta["sizes"]))
        kg.relations = {m: [tuple(r) for r in v] for m, v in meta["relations"].items()}
        kg.rel_keys = [(m1, r[1], r[0]) for m1 in kg.relations for r in kg.relations[m1]]
        kg.out = {m: [k for k in kg.rel_keys if k[0] == m] for m in kg.modes}
        kg.node_ids = {m: z["ids/" + m] for m in kg.modes}
        t = lambda a: torch.from_numpy(np.array(a))
        tables = {m: t(z["table/" + m]) for m in kg.modes}
        pre = {m: t(z["pre/" + m]) for m in kg.modes if "pre/" + m in z.files}
        post = {m: t(z["post/" + m]) for m in kg.modes if "post/" + m in z.files}
        rel_params = {rel: t(z["rel/%d" % i]) for i, rel in enumerate(kg.rel_keys)}
        batches = {}
        for s, rels in meta["rels"].items():
            batches[s] = {"rels": _tup(rels), "target": z["q/%s/target" % s], "anchors": z["q/%s/anchors" % s],
                          "negs": z["q/%s/negs" % s]}
        case = Case(kg, meta["d"], meta["decoder"], meta["inter"], tables, rel_params, pre, post, batches)
        expected = {k[4:]: z[k] for k in z.files if k.startswith("exp/")}
        expected.update(meta.get("exp", {}))
        return case, expected


def make_case(seed, d, decoder, inter, n_modes=3, nodes_per_mode=40, n_rel_pairs=4, n_queries=24, n_neg=5,
              structures=STRUCTURES, kg=None):
    """Draw a case.  Small by default (golden fixtures); the GPU tests call it
    with larger sizes."""
    rng = np.random.RandomState(seed)
    gen = torch.Generator().manual_seed(seed)
    if kg is None:
        kg = SynthKG(["m%d" % i for i in range(n_modes)], [nodes_per_mode] * n_modes, n_rel_pairs, seed=seed)
    # N_mode + 2 rows: positions 0..N-1 shift to rows 1..N; row 0 and row N+1 exist
    # but are never addressed by real nodes (bio/data_utils.py:13-16)
    tables = {m: torch.randn(kg.sizes[m] + 2, d, generator=gen) * (1.0 / d) for m in kg.modes}
    rel_params = {}
    for rel in kg.rel_keys:
        if decoder == "bilinear":
            bound = math.sqrt(6.0 / (2 * d))
            rel_params[rel] = (torch.rand(d, d, generator=gen) * 2 - 1) * bound
        else:
            bound = 6.0 / math.sqrt(d)
            rel_params[rel] = (torch.rand(d, generator=gen) * 2 - 1) * bound
    pre, post = {}, {}
    if not inter.endswith("-simple"):
        bound = math.sqrt(6.0 / (2 * d))
        for m in kg.modes:
            pre[m] = (torch.rand(d, d, generator=gen) * 2 - 1) * bound
            post[m] = (torch.rand(d, d, generator=gen) * 2 - 1) * bound
    batches = {}
    for s in structures:
        rels = kg.sample_rels(s, rng)
        b = kg.sample_batch(s, rels, n_queries, n_neg, rng)
        b["rels"] = rels
        batches[s] = b
    return Case(kg, d, decoder, inter, tables, rel_params, pre, post, batches)
