"""Shared set-up of the GraphSAGE-style encoder tests (f4): a small typed graph with adjacency
lists (some nodes without neighbours, some with more than max_keep), tables with the reference's
null row, compression matrices per layer."""
import numpy as np
import torch

from graphqembed_b200.synth import SynthKG


def make_sage_case(seed=0, d=32, nodes_per_mode=60, n_modes=3, n_rel_pairs=5):
    rng = np.random.RandomState(seed)
    kg = SynthKG(["m%d" % i for i in range(n_modes)], [nodes_per_mode] * n_modes, n_rel_pairs, seed=seed)
    node_maps = kg.node_maps()
    for m in node_maps:
        node_maps[m][-1] = -1                       # the null neighbour (bio/data_utils.py:14-15) -> row 0
    gen = torch.Generator().manual_seed(seed)
    tables = {m: torch.randn(len(node_maps[m]) + 1, d, generator=gen) / d for m in kg.modes}
    adj = {}
    for rel in kg.rel_keys:
        src, dst = kg.node_ids[rel[0]], kg.node_ids[rel[2]]
        adj[rel] = {}
        for n in src:
            deg = int(rng.choice([0, 1, 2, 5, 30]))
            adj[rel][int(n)] = [int(x) for x in rng.choice(dst, size=deg, replace=False)]
    n_in = {m: d * (1 + len(kg.relations[m])) for m in kg.modes}
    compress = [{m: (torch.rand(d, n_in[m], generator=gen) - 0.5) * 0.6 for m in kg.modes} for _ in range(3)]
    return kg, node_maps, tables, adj, compress
