"""f4: the oracle's GraphSAGE-style encoder against the UNMODIFIED reference
``Encoder`` / ``MeanAggregator`` (netquery/encoders.py:47-129, netquery/aggregators.py:17-68),
stacked as netquery/utils.py:93-126 does, same ``random`` seed -> bit-exact."""
import random

import pytest
import torch

from oracle import netquery_oracle as O
from oracle import ref_shim
from sage_common import make_sage_case

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")


def _reference_stack(depth, tables, node_maps, kg, adj, compress):
    _, _, _, E = ref_shim.load()
    import netquery.aggregators as A
    d = next(iter(tables.values())).size(1)
    fm = ref_shim.IterDict()
    for m in kg.modes:
        emb = torch.nn.Embedding(tables[m].size(0), d)
        emb.weight.data.copy_(tables[m])
        fm[m] = emb
    features = lambda nodes, mode: fm[mode](torch.LongTensor([node_maps[mode][n] for n in nodes]) + 1)
    dims = ref_shim.IterDict({m: d for m in kg.modes})

    def layer(k, feats, agg_feats, **kw):
        enc = E.Encoder(feats, dims, dims, kg.relations, adj, aggregator=A.MeanAggregator(agg_feats), **kw)
        for m in kg.modes:
            enc.compress_params[m].data.copy_(compress[k][m])
        return enc

    # netquery/utils.py:103-126 (the module itself needs py2's cPickle, so its wiring is restated)
    enc1 = layer(0, features, features, feature_modules=fm)
    if depth == 1:
        return enc1
    lower1 = lambda nodes, mode: enc1(nodes, mode).t().squeeze()
    enc2 = layer(1, lower1, lower1, base_model=enc1, feature_modules=ref_shim.IterDict())   # (py2 dict default)
    if depth == 2:
        return enc2
    lower2 = lambda nodes, mode: enc2(nodes, mode).t().squeeze()
    return layer(2, lower1, lower2, base_model=enc2, feature_modules=ref_shim.IterDict())


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_sage_oracle_bit_exact_vs_reference(depth):
    kg, node_maps, tables, adj, compress = make_sage_case(seed=depth)
    ref = _reference_stack(depth, tables, node_maps, kg, adj, compress)
    orc_feats = O.OracleScorer(tables, node_maps, {}, "bilinear", "mean-simple").raw_features
    orc = O.sage_stack(depth, orc_feats, kg.relations, adj, compress)
    for mode in kg.modes:
        nodes = [int(n) for n in kg.node_ids[mode][:17]] + [-1]
        random.seed(99)
        want = ref(nodes, mode)
        random.seed(99)
        got = orc(nodes, mode)
        assert want.shape == (32, len(nodes))
        assert torch.equal(want.detach(), got), (depth, mode)
