"""Synthetic, seeded, Bio-shaped knowledge graphs and query batches.

The Bio KG itself is an external download (reference README.md:15) and there
is no network, so every test and benchmark runs on a synthetic graph of the
same shape: a handful of node types ("modes"), typed directed relations
closed under reversal (reference netquery/graph.py:4-5 assumes both
directions exist), node ids that are arbitrary integers mapped to table rows
through ``node_maps`` (reference netquery/bio/data_utils.py:13-21).

Only index data is produced here (which node feeds which operand); parameters
are created by the operator classes themselves.
"""
import numpy as np

STRUCTURES = ("1-chain", "2-chain", "3-chain", "2-inter", "3-inter",
              "3-inter_chain", "3-chain_inter")
N_ANCHORS = {"1-chain": 1, "2-chain": 1, "3-chain": 1, "2-inter": 2,
             "3-inter": 3, "3-inter_chain": 2, "3-chain_inter": 2}

BIO_MODES = ("protein", "function", "disease", "drug", "sideeffect")
BIO_SIZES = (40000, 20000, 10000, 7000, 20000)   # 97 000 nodes (arbitrary split)


class SynthKG(object):
    """A typed relation schema plus per-mode node id lists."""

    def __init__(self, modes, sizes, n_rel_pairs, seed=0, self_loops=True):
        rng = np.random.RandomState(seed)
        self.modes = list(modes)
        self.sizes = {m: int(s) for m, s in zip(modes, sizes)}
        # Global integer node ids; each mode owns a contiguous id range but the
        # id -> row mapping inside the mode is a random permutation, so that the
        # node_maps indirection of the reference is really exercised.
        self.node_ids = {}
        base = 0
        for m in self.modes:
            self.node_ids[m] = base + rng.permutation(self.sizes[m]).astype(np.int64)
            base += self.sizes[m]
        # Relations: unordered mode pairs -> two directed triples each; a pair
        # (m, m) yields one triple that is its own reverse.
        self.relations = {m: [] for m in self.modes}
        pairs = []
        nm = len(self.modes)
        # a ring first, so that every mode has outgoing relations
        for i in range(nm):
            pairs.append((i, (i + 1) % nm))
        while len(pairs) < n_rel_pairs:
            i, j = int(rng.randint(nm)), int(rng.randint(nm))
            if i == j and not self_loops:
                continue
            pairs.append((i, j))
        for k, (i, j) in enumerate(pairs[:n_rel_pairs]):
            name = "r%d" % k
            a, b = self.modes[i], self.modes[j]
            self.relations[a].append((b, name))
            if a != b:
                self.relations[b].append((a, name))
        # canonical triples in the decoders' registration order
        # (reference netquery/decoders.py:135-137)
        self.rel_keys = [(m1, r[1], r[0]) for m1 in self.relations for r in self.relations[m1]]
        self.out = {m: [k for k in self.rel_keys if k[0] == m] for m in self.modes}

    # ---- id <-> row ------------------------------------------------------
    def node_maps(self):
        """{mode: {node_id: position}} as bio/data_utils.py:13 builds it."""
        return {m: {int(n): i for i, n in enumerate(ids)} for m, ids in self.node_ids.items()}

    def full_lists(self):
        return {m: [int(n) for n in ids] for m, ids in self.node_ids.items()}

    # ---- formulas -----------------------------------------------------------
    def sample_rels(self, structure, rng, target_modes=None):
        """Type-consistent relation tuple for ``structure`` (shape of
        ``Formula.rels``, reference netquery/graph.py:40-53).  ``target_modes``
        restricts the target node type (queries routed to the rank that owns it)."""
        pick = lambda m: self.out[m][int(rng.randint(len(self.out[m])))]
        pool = self.modes if target_modes is None else list(target_modes)
        t = pool[int(rng.randint(len(pool)))]
        if structure.endswith("-chain") and structure[0] in "123":
            rels, m = [], t
            for _ in range(int(structure[0])):
                r = pick(m)
                rels.append(r)
                m = r[2]
            return tuple(rels)
        if structure in ("2-inter", "3-inter"):
            return tuple(pick(t) for _ in range(int(structure[0])))
        if structure == "3-inter_chain":
            r1, r2a = pick(t), pick(t)
            return (r1, (r2a, pick(r2a[2])))
        if structure == "3-chain_inter":
            r1 = pick(t)
            return (r1, (pick(r1[2]), pick(r1[2])))
        raise ValueError(structure)

    @staticmethod
    def modes_of(structure, rels):
        """(target_mode, anchor_modes) -- reference netquery/graph.py:15-24."""
        t = rels[0][0]
        if structure in ("1-chain", "2-chain", "3-chain"):
            return t, (rels[-1][-1],)
        if structure in ("2-inter", "3-inter"):
            return t, tuple(r[-1] for r in rels)
        if structure == "3-inter_chain":
            return t, (rels[0][-1], rels[1][-1][-1])
        return t, (rels[1][0][-1], rels[1][1][-1])

    # ---- batches ------------------------------------------------------------
    def sample_nodes(self, mode, n, rng):
        ids = self.node_ids[mode]
        return ids[rng.randint(0, len(ids), size=n)]

    def sample_batch(self, structure, rels, n_queries, n_neg, rng):
        """Uniform node ids: dict(target [B], anchors [A,B], negs [B,K])."""
        t, amodes = self.modes_of(structure, rels)
        return {
            "target": self.sample_nodes(t, n_queries, rng),
            "anchors": np.stack([self.sample_nodes(m, n_queries, rng) for m in amodes]),
            "negs": self.sample_nodes(t, n_queries * n_neg, rng).reshape(n_queries, n_neg),
        }

    @staticmethod
    def query_graph(structure, rels, target, anchors, mid=-7):
        """The nested-tuple encoding Query.__init__ consumes (reference
        netquery/graph.py:40-54).  Intermediate (unobserved) nodes get ``mid``."""
        t, a = int(target), [int(x) for x in anchors]
        if structure == "1-chain":
            return (structure, (t, rels[0], a[0]))
        if structure == "2-chain":
            return (structure, (t, rels[0], mid), (mid, rels[1], a[0]))
        if structure == "3-chain":
            return (structure, (t, rels[0], mid), (mid, rels[1], mid - 1), (mid - 1, rels[2], a[0]))
        if structure in ("2-inter", "3-inter"):
            return (structure,) + tuple((t, r, x) for r, x in zip(rels, a))
        if structure == "3-inter_chain":
            return (structure, (t, rels[0], a[0]),
                    ((t, rels[1][0], mid), (mid, rels[1][1], a[1])))
        if structure == "3-chain_inter":
            return (structure, (t, rels[0], mid),
                    ((mid, rels[1][0], a[0]), (mid, rels[1][1], a[1])))
        raise ValueError(structure)


def bio_shaped(seed=0, scale=1.0):
    """5 modes / 97 000 nodes / 42 directed relation triples."""
    sizes = [max(8, int(s * scale)) for s in BIO_SIZES]
    # 21 pairs of distinct modes -> 42 directed triples
    return SynthKG(BIO_MODES, sizes, n_rel_pairs=21, seed=seed, self_loops=False)


def synthetic_large(n_modes=8, nodes_per_mode=1250000, n_rel_pairs=50, seed=0):
    """Config 5 shape: 10 M nodes / 100 directed relations / 8 modes."""
    modes = ["m%d" % i for i in range(n_modes)]
    return SynthKG(modes, [nodes_per_mode] * n_modes, n_rel_pairs, seed=seed, self_loops=False)
