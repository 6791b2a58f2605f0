"""f4 on the GPU: the package's GraphSAGE-style encoder (``graphqembed_b200/sage.py`` over
``gqe_segment_mean_device`` / ``gqe_linear_device``) against the oracle restatement, which
``tests/test_sage_oracle.py`` pins bit-exact to the reference.  Same ``random`` seed -> same
sampled neighbours; values within 1e-5 (fp32, different summation order)."""
import random

import numpy as np
import pytest
import torch

from oracle import netquery_oracle as O
from sage_common import make_sage_case

pytestmark = pytest.mark.gpu


class _Graph(object):
    def __init__(self, kg, lookup, adj, d):
        self.features, self.relations, self.adj_lists = lookup, kg.relations, adj
        self.feature_dims = {m: d for m in kg.modes}
        self.full_lists = kg.full_lists()


def _package_encoder(depth, kg, node_maps, tables, adj, compress, d):
    import graphqembed_b200 as gqe
    fm = {}
    for m in kg.modes:
        emb = torch.nn.Embedding(tables[m].size(0), d)
        emb.weight.data.copy_(tables[m])
        fm[m] = emb
    graph = _Graph(kg, gqe.RowLookup(node_maps), adj, d)
    enc = gqe.get_encoder(depth, graph, {m: d for m in kg.modes}, fm)
    layers, e = [], enc
    while e is not None:
        layers.append(e)
        e = getattr(e, "base_model", None)
    for k, layer in enumerate(reversed(layers)):
        for m in kg.modes:
            layer.compress_params[m].data.copy_(compress[k][m])
    return enc.to("cuda"), graph


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_sage_encoder_matches_oracle(depth):
    d = 32
    kg, node_maps, tables, adj, compress = make_sage_case(seed=depth, d=d)
    enc, _ = _package_encoder(depth, kg, node_maps, tables, adj, compress, d)
    feats = O.OracleScorer(tables, node_maps, {}, "bilinear", "mean-simple").raw_features
    orc = O.sage_stack(depth, feats, kg.relations, adj, compress)
    for mode in kg.modes:
        nodes = [int(n) for n in kg.node_ids[mode][:23]] + [-1]
        random.seed(5)
        want = orc(nodes, mode)
        random.seed(5)
        got = enc(nodes, mode)
        assert got.is_cuda and got.shape == want.shape
        assert float((got.cpu() - want).abs().max()) < 1e-5, (depth, mode)
    names = set(enc.state_dict().keys())
    assert all("%s_compress" % m in names for m in kg.modes)           # encoders.py:102


def test_segment_mean_and_linear_calls():
    """The two C-ABI calls on their own: ragged segments (1 .. 40 rows), d = 256; a non-square
    compression with ReLU."""
    import graphqembed_b200 as gqe
    rng = np.random.RandomState(0)
    ctx = gqe.Context(0)
    src = torch.randn(500, 256, device="cuda")
    lens = rng.randint(1, 41, size=300)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cols = rng.randint(0, 500, size=int(ptr[-1])).astype(np.int32)
    out = torch.empty(300, 256, device="cuda")
    d_ptr, d_cols = torch.from_numpy(ptr).cuda(), torch.from_numpy(cols).cuda()     # (kept alive across the launch)
    ctx.segment_mean_device(src.data_ptr(), 500, 256, 300, d_ptr.data_ptr(), d_cols.data_ptr(), out.data_ptr())
    want = torch.stack([src[torch.from_numpy(cols[ptr[i]:ptr[i + 1]]).long().cuda()].mean(0) for i in range(300)])
    assert float((out - want).abs().max()) < 1e-6
    w, x = torch.randn(96, 200, device="cuda"), torch.randn(200, 77, device="cuda")
    y = torch.empty(96, 77, device="cuda")
    ctx.linear_device(w.data_ptr(), 96, 200, 77, x.data_ptr(), 1, y.data_ptr())
    torch.cuda.synchronize()
    assert float((y - torch.relu(w.double().mm(x.double())).float()).abs().max()) < 1e-4
    # a column outside the source is reported like every other bad row index
    bad = torch.tensor([0, 700], dtype=torch.int32, device="cuda")
    one = torch.tensor([0, 2], dtype=torch.int64, device="cuda")
    ctx.segment_mean_device(src.data_ptr(), 500, 256, 1, one.data_ptr(), bad.data_ptr(), out.data_ptr())
    with pytest.raises(IndexError):
        ctx.index_error()


@pytest.mark.parametrize("decoder,inter", [("bilinear", "mean"), ("transe", "min-simple")])
def test_query_scores_with_sage_encoder(decoder, inter):
    """QueryEncoderDecoder over a depth-2 encoder: every structure, scores and margin loss
    against the oracle's model.py:70-127 with the oracle encoder plugged in."""
    import graphqembed_b200 as gqe
    from oracle.cases import make_case
    d = 32
    kg, node_maps, tables, adj, compress = make_sage_case(seed=7, d=d, nodes_per_mode=80)
    case = make_case(seed=3, d=d, decoder=decoder, inter=inter, n_queries=24, n_neg=2, kg=kg)
    enc, graph = _package_encoder(2, kg, node_maps, tables, adj, compress, d)
    dims = {m: d for m in kg.modes}
    dec = gqe.get_metapath_decoder(graph, dims, decoder)
    store = dec.mats if decoder == "bilinear" else dec.vecs
    for rel, p in case.rel_params.items():
        store[rel].data.copy_(p)
    idec = gqe.get_intersection_decoder(graph, dims, inter)
    if not inter.endswith("-simple"):
        for m in kg.modes:
            idec.pre_mats[m].data.copy_(case.pre[m])
            idec.post_mats[m].data.copy_(case.post[m])
    model = gqe.QueryEncoderDecoder(graph, enc, dec, idec).to("cuda")
    orc = O.OracleScorer(tables, node_maps, case.rel_params, decoder, inter, case.pre, case.post,
                         full_lists=kg.full_lists())
    orc.encoder = O.sage_stack(2, orc.raw_features, kg.relations, adj, compress)
    for s in case.batches:
        f_o, f_p = case.formula(s), case.formula(s, cls=gqe.Formula)
        q_o, q_p = case.queries(s), case.queries(s, cls=gqe.Query)
        targets = [int(t) for t in case.batches[s]["target"]]
        random.seed(11)
        want = orc.forward(f_o, q_o, targets)
        random.seed(11)
        got = model.forward(f_p, q_p, targets)
        assert float((got.cpu() - want).abs().max()) < 1e-4, s
        random.seed(12)
        want_loss = orc.margin_loss(f_o, q_o)
        random.seed(12)
        got_loss = model.margin_loss(f_p, q_p)
        assert abs(float(got_loss) - float(want_loss)) < 1e-4, s
    with pytest.raises(RuntimeError):
        model.context()
