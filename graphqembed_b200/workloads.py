"""BASELINE.json workloads as synthetic, seeded index batches.

A workload is a list of (Formula, node-id arrays) slices plus the graph they
were drawn from; ``lower()`` turns it into the flat int32 row arrays and the
``gqe_segment`` table the grouped C-ABI entry points take.  Parameters are NOT
part of a workload (they belong to the operator modules / the raw device
buffers of the caller).
"""
import numpy as np

from . import _lib
from .lowering import RowLookup, lower_formula
from .query import Formula, QueryBatch
from .synth import N_ANCHORS, STRUCTURES, bio_shaped, synthetic_large

# name -> (description, d, total queries, structures, decoder, inter)
WORKLOADS = {
    # BASELINE.json configs[1]
    "bio-chain-d128-b4096": ("Bio KG 2/3-chain path queries, Bilinear decoder, d=128, batch=4096", 128, 4096,
                             ("2-chain", "3-chain"), "bilinear", "mean"),
    # BASELINE.json configs[2]
    "bio-inter-d128-b8192": ("Bio KG 2-inter + 3-inter DeepSets intersection, d=128, batch=8192", 128, 8192,
                             ("2-inter", "3-inter"), "bilinear", "mean"),
    # BASELINE.json configs[3] -- the largest single-GPU configuration: the bench default
    "bio-mix-d256-b65536": ("Bio KG full mix (1/2/3-chain + 2/3-inter + 3-inter_chain), d=256, batch=65536", 256,
                            65536, STRUCTURES[:6], "bilinear", "mean"),
    # configs[2] with the element-wise-min aggregator the north star names (utils.py:144-145)
    "bio-inter-d128-b8192-min": ("Bio KG 2-inter + 3-inter DeepSets intersection (agg = min), d=128, batch=8192", 128,
                                 8192, ("2-inter", "3-inter"), "bilinear", "min"),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "bio-edge-d128-b512": ("Bio KG 1-chain (edge) queries, Bilinear decoder, d=128, batch=512", 128, 512,
                           ("1-chain",), "bilinear", "mean"),
    # BASELINE.json configs[4]: 8 node types x 1.25 M nodes, 100 directed relations; the table (10.24 GB)
    # is sharded by node type over the GPUs of one box (1.28 GB per GPU at 8 GPUs)
    "synth-10m-d256-b65536": ("Synthetic KG 10M nodes / 100 relations, full query mix, d=256, batch=65536 per GPU",
                              256, 65536, STRUCTURES[:6], "bilinear", "mean"),
    # the contraction-free operators (TransE / DistMult decoder, SimpleSetIntersection): pure HBM gather.
    # On the 10 M-node KG the table (10.24 GB) is far beyond L2, so every row comes from DRAM.
    "synth-10m-transe-minsimple-d256-b65536": ("Synthetic KG 10M nodes, full query mix, TransE decoder + "
                                               "SimpleSetIntersection(min), d=256, batch=65536", 256, 65536,
                                               STRUCTURES[:6], "transe", "min-simple"),
    "synth-10m-distmult-meansimple-d256-b65536": ("Synthetic KG 10M nodes, full query mix, DistMult decoder + "
                                                  "SimpleSetIntersection(mean), d=256, batch=65536", 256, 65536,
                                                  STRUCTURES[:6], "bilinear-diag", "mean-simple"),
    "bio-transe-minsimple-d256-b65536": ("Bio KG full query mix, TransE decoder + SimpleSetIntersection(min), d=256, "
                                         "batch=65536 (tables L2-resident)", 256, 65536, STRUCTURES[:6], "transe",
                                         "min-simple"),
}
DEFAULT_WORKLOAD = "bio-mix-d256-b65536"
LARGE_WORKLOAD = "synth-10m-d256-b65536"


class Workload(object):
    def __init__(self, name, kg, d, decoder, inter, batches):
        self.name, self.kg, self.d, self.decoder, self.inter = name, kg, d, decoder, inter
        self.batches = batches                      # [QueryBatch] with (positive, negative) per query
        self.n_queries = sum(b.n_queries for b in batches)

    def lower(self, lookup, mode_ids, rel_ids):
        """-> (segments ctypes array, anchor_rows int32 [3, Q], pair_rows int32 [Q, 2])"""
        total = self.n_queries
        anchor_rows = np.zeros((_lib.GQE_MAX_ANCHORS, total), dtype=np.int32)
        pair_rows = np.empty((total, 2), dtype=np.int32)
        items, q0 = [], 0
        for b in self.batches:
            f = b.formula
            for k, mode in enumerate(f.anchor_modes):
                anchor_rows[k, q0:q0 + b.n_queries] = lookup.rows(b.anchors[k], mode)
            pair_rows[q0:q0 + b.n_queries] = lookup.rows(b.targets, f.target_mode).reshape(-1, 2)
            items.append((lower_formula(f, mode_ids, rel_ids), q0, q0 + b.n_queries))
            q0 += b.n_queries
        return _lib.make_segments(items), anchor_rows, pair_rows

    def node_arrays(self, mode_ids, rel_ids):
        """-> (segments, anchor_nodes int32 [3, Q], pair_nodes int32 [Q, 2]): the reference's NODE IDS
        in the layout of the grouped *_nodes entry points (no lookup done here)."""
        total = self.n_queries
        anchor_nodes = np.zeros((_lib.GQE_MAX_ANCHORS, total), dtype=np.int32)
        pair_nodes = np.empty((total, 2), dtype=np.int32)
        items, q0 = [], 0
        for b in self.batches:
            anchor_nodes[:b.anchors.shape[0], q0:q0 + b.n_queries] = b.anchors
            pair_nodes[q0:q0 + b.n_queries] = b.targets.reshape(-1, 2)
            items.append((lower_formula(b.formula, mode_ids, rel_ids), q0, q0 + b.n_queries))
            q0 += b.n_queries
        return _lib.make_segments(items), anchor_nodes, pair_nodes

    # SURVEY.md section 8d, per query, fp32 rows + int32 indices + fp32 scores, T = 2
    def algorithmic_bytes(self):
        total = 0
        for b in self.batches:
            a = len(b.formula.anchor_modes)
            total += b.n_queries * ((a + 2) * 4 * self.d + (a + 2) * 4 + 2 * 4)
        return total

    def algorithmic_flops(self):
        """Bilinear + DeepSets contractions only, query side counted once."""
        if self.decoder != "bilinear":
            return 0
        d2 = self.d * self.d
        simple = self.inter.endswith("-simple")
        per = {"1-chain": 2 * 2 * d2, "2-chain": 2 * 4 * d2, "3-chain": 2 * 6 * d2,
               "2-inter": (4 if simple else 10) * d2, "3-inter": (6 if simple else 14) * d2,
               "3-inter_chain": (6 if simple else 12) * d2, "3-chain_inter": (6 if simple else 12) * d2}
        return sum(b.n_queries * per[b.formula.query_type] for b in self.batches)


    def composed_flops(self, min_rows=1024):
        """Contraction flops the tensor-core path actually runs once runs of linear operators
        are pre-multiplied (gqe_compose; formulas with >= min_rows tile rows): one contraction
        per (query, target) pair for chains, one per branch + post for DeepSets intersections."""
        if self.decoder != "bilinear":
            return 0
        d2 = self.d * self.d
        simple = self.inter.endswith("-simple")
        total = 0
        for b in self.batches:
            qt, nq = b.formula.query_type, b.n_queries
            chain = qt.endswith("-chain") and qt[0] in "123"
            rows = 2 * nq if chain else nq
            if rows < min_rows:
                per = {"1-chain": 4, "2-chain": 8, "3-chain": 12, "2-inter": 4 if simple else 10,
                       "3-inter": 6 if simple else 14, "3-inter_chain": 6 if simple else 12,
                       "3-chain_inter": 6 if simple else 12}[qt]
            elif chain:
                per = 4
            elif simple:
                per = {"2-inter": 4, "3-inter": 6, "3-inter_chain": 4, "3-chain_inter": 6}[qt]
            else:
                per = {"2-inter": 6, "3-inter": 8, "3-inter_chain": 6, "3-chain_inter": 6}[qt]
            total += nq * per * d2
        return total


_KG_CACHE = {}


def _kg_cached(kind):
    kg = _KG_CACHE.get(kind)
    if kg is None:
        kg = _KG_CACHE[kind] = synthetic_large(seed=0) if kind == "large" else bio_shaped(seed=0)
    return kg


def make_workload(name=DEFAULT_WORKLOAD, seed=0, kg=None, formulas_per_structure=1, total=None, target_modes=None):
    desc, d, n_total, structures, decoder, inter = WORKLOADS[name]
    if total is not None:
        n_total = int(total)
    if kg is None:
        kg = _kg_cached("large" if name.startswith("synth-10m") else "bio")
    rng = np.random.RandomState(1000 + seed)
    n_slices = len(structures) * formulas_per_structure
    sizes = [n_total // n_slices + (1 if i < n_total % n_slices else 0) for i in range(n_slices)]
    batches, i = [], 0
    for s in structures:
        for _ in range(formulas_per_structure):
            rels = kg.sample_rels(s, rng, target_modes)
            b = kg.sample_batch(s, rels, sizes[i], 1, rng)
            pairs = np.stack([b["target"], b["negs"][:, 0]], axis=1).reshape(-1)
            batches.append(QueryBatch(Formula(s, rels), b["anchors"], pairs))
            i += 1
    return Workload(name, kg, d, decoder, inter, batches)
