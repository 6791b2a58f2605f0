"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the conjunctive-query scoring path.

A restatement, in this repo's own words, of what the reference computes on the
path ``QueryEncoderDecoder.forward`` / ``margin_loss`` (reference
``netquery/model.py:70-127``).  All arithmetic is delegated to the same ATen
CPU operators the reference calls (``mm``, embedding lookup, ``norm``, ``div``,
``relu``, ``stack``, ``mean``/``min``, ``cosine_similarity``) in the same
order, so on identical parameters and inputs the oracle reproduces the
reference bit for bit (checked by ``tests/test_oracle_vs_reference.py`` when
``/root/reference`` is mounted, and frozen in ``tests/golden/*.npz``).

Parity status: the reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF run in the build container (``oracle/ref_shim.py`` +
``oracle/make_golden.py``), not against reference-authored known answers.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.  The product (``graphqembed_b200``) never does.
"""
import math
import random

import torch
import torch.nn.functional as F

CHAIN_TYPES = ("1-chain", "2-chain", "3-chain")
FLAT_INTER_TYPES = ("2-inter", "3-inter")
ALL_TYPES = CHAIN_TYPES + FLAT_INTER_TYPES + ("3-inter_chain", "3-chain_inter")


def reverse_relation(rel):
    """(m1, name, m2) -> (m2, name, m1).  Reference: netquery/graph.py:4-5."""
    return (rel[-1], rel[1], rel[0])


class Formula(object):
    """Query structure + typed relations.  Reference: netquery/graph.py:11-36.

    ``rels`` is a tuple of relation triples; for the two nested structures the
    second entry is itself a pair of triples.  ``target_mode`` is the source
    mode of the first relation; ``anchor_modes`` follow graph.py:17-24.
    """

    def __init__(self, query_type, rels):
        self.query_type = query_type
        self.rels = rels
        self.target_mode = rels[0][0]
        if query_type in CHAIN_TYPES:
            self.anchor_modes = (rels[-1][-1],)
        elif query_type in FLAT_INTER_TYPES:
            self.anchor_modes = tuple(r[-1] for r in rels)
        elif query_type == "3-inter_chain":
            self.anchor_modes = (rels[0][-1], rels[1][-1][-1])
        elif query_type == "3-chain_inter":
            self.anchor_modes = (rels[1][0][-1], rels[1][1][-1])

    def _key(self):
        return (self.query_type, self.rels)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return self._key() == other._key()

    def __str__(self):
        return "%s: %s" % (self.query_type, self.rels)


class Query(object):
    """One sampled query.  Reference: netquery/graph.py:38-66.

    ``query_graph`` = (type, edge, edge | (edge, edge), ...), an edge being
    ``(node_u, (mode_u, rel, mode_v), node_v)``.  The target node is always
    ``query_graph[1][0]`` (graph.py:54); anchors per structure follow
    graph.py:42-53.
    """

    def __init__(self, query_graph, neg_samples, hard_neg_samples, neg_sample_max=100):
        qt = query_graph[0]
        edges = query_graph[1:]
        if qt in CHAIN_TYPES:
            self.formula = Formula(qt, tuple(e[1] for e in edges))
            self.anchor_nodes = (edges[-1][-1],)
        elif qt in FLAT_INTER_TYPES:
            self.formula = Formula(qt, tuple(e[1] for e in edges))
            self.anchor_nodes = tuple(e[-1] for e in edges)
        elif qt == "3-inter_chain":
            self.formula = Formula(qt, (edges[0][1], (edges[1][0][1], edges[1][1][1])))
            self.anchor_nodes = (edges[0][-1], edges[1][-1][-1])
        elif qt == "3-chain_inter":
            self.formula = Formula(qt, (edges[0][1], (edges[1][0][1], edges[1][1][1])))
            self.anchor_nodes = (edges[1][0][-1], edges[1][1][-1])
        self.target_node = edges[0][0]
        # graph.py:59-66: lists are truncated by random.sample when too long
        # (note the reference's asymmetric < vs <= between the two lists).
        if neg_samples is None:
            self.neg_samples = None
        elif len(neg_samples) < neg_sample_max:
            self.neg_samples = list(neg_samples)
        else:
            self.neg_samples = random.sample(list(neg_samples), neg_sample_max)
        if hard_neg_samples is None:
            self.hard_neg_samples = None
        elif len(hard_neg_samples) <= neg_sample_max:
            self.hard_neg_samples = list(hard_neg_samples)
        else:
            self.hard_neg_samples = random.sample(list(hard_neg_samples), neg_sample_max)


def _amin(x, dim):
    # The reference passes torch.min and unwraps a tuple (decoders.py:296-298);
    # on torch>=1.x the returned named tuple fails its ``type(..) == tuple``
    # test, so the oracle (and ref_shim) use the values directly.
    return torch.min(x, dim=dim)[0]


class OracleScorer(object):
    """The reference's DirectEncoder + metapath decoder + intersection + cosine.

    tables      {mode: FloatTensor[rows, d]}; row = node_maps[mode][node] + 1
                (netquery/bio/data_utils.py:13-21).  node_maps may be None, in
                which case the node id is the 0-based index (utils.py:18-20).
    rel_params  {(m1, name, m2): FloatTensor[d, d]} for 'bilinear'
                (decoders.py:139), FloatTensor[d] for 'transe' /
                'bilinear-diag' (decoders.py:195,224).
    decoder     'bilinear' | 'transe' | 'bilinear-diag'   (utils.py:128-137)
    inter       'mean' | 'min' | 'mean-simple' | 'min-simple' (utils.py:139-150)
    pre, post   {mode: FloatTensor[d_exp, d]}, {mode: FloatTensor[d, d_exp]}
                (decoders.py:282-286); unused for the -simple kinds.
    full_lists  {mode: [node, ...]} -- only for 1-chain negatives (model.py:118)
    """

    def __init__(self, tables, node_maps, rel_params, decoder, inter,
                 pre=None, post=None, full_lists=None, dtype=torch.float32):
        cast = lambda t: t.detach().to("cpu", dtype).contiguous()
        self.tables = {m: cast(t) for m, t in tables.items()}
        self.node_maps = node_maps
        self.rel_params = {r: cast(t) for r, t in rel_params.items()}
        self.decoder = decoder
        self.inter = inter
        self.pre = {m: cast(t) for m, t in (pre or {}).items()}
        self.post = {m: cast(t) for m, t in (post or {}).items()}
        self.full_lists = full_lists
        self.trace = None  # when a list: receives ("rows", mode, [...]) / ("rel", key)
        self.encoder = None  # a callable (nodes, mode) -> [d, B] replacing the DirectEncoder (OracleSageEncoder)

    # ---- a4/a5: features closure + DirectEncoder ---------------------------
    def rows_of(self, nodes, mode):
        """bio/data_utils.py:20-21: node id -> node_maps[mode][n] -> +1."""
        if self.node_maps is None:
            idx = torch.LongTensor(list(nodes)) + 1
        else:
            nm = self.node_maps[mode]
            idx = torch.LongTensor([nm[n] for n in nodes]) + 1
        if self.trace is not None:
            self.trace.append(("rows", mode, idx.tolist()))
        return idx

    def encode(self, nodes, mode):
        """encoders.py:41-43: lookup, transpose to [d, B], divide by the column
        L2 norm (no epsilon: an all-zero row yields NaN, as in the reference)."""
        if self.encoder is not None:
            return self.encoder(nodes, mode)
        embeds = F.embedding(self.rows_of(nodes, mode), self.tables[mode]).t()
        norm = embeds.norm(p=2, dim=0, keepdim=True)
        return embeds.div(norm.expand_as(embeds))

    def raw_features(self, nodes, mode):
        """The ``features`` closure itself (bio/data_utils.py:20-21): raw table rows [n, d]."""
        return F.embedding(self.rows_of(nodes, mode), self.tables[mode])

    def _param(self, rel):
        if self.trace is not None:
            self.trace.append(("rel", rel))
        return self.rel_params[rel]

    # ---- a6-a9: metapath decoders -----------------------------------------
    def project(self, embeds, rel):
        p = self._param(rel)
        if self.decoder == "bilinear":          # decoders.py:149-150
            return p.mm(embeds)
        col = p.unsqueeze(1).expand(p.size(0), embeds.size(1))
        if self.decoder == "transe":            # decoders.py:207-208
            return embeds + col
        return embeds * col                      # decoders.py:235-236

    def path_score(self, embeds1, embeds2, rels):
        if self.decoder == "bilinear":          # decoders.py:142-147
            act = embeds1.t()
            for r in rels:
                act = act.mm(self._param(r))
            return F.cosine_similarity(act.t(), embeds2, dim=0, eps=1e-8)
        if self.decoder == "transe":            # decoders.py:200-205 (in place)
            moved = embeds1
            for r in rels:
                v = self._param(r)
                moved += v.unsqueeze(1).expand(v.size(0), embeds1.size(1))
            return F.cosine_similarity(embeds2, moved, dim=0, eps=1e-8)
        acts = embeds1                           # decoders.py:228-233 (raw dot)
        for r in rels:
            v = self._param(r)
            acts = acts * v.unsqueeze(1).expand(v.size(0), embeds1.size(1))
        return (acts * embeds2).sum(0)

    # ---- a10/a11: intersections --------------------------------------------
    def intersect(self, e1, e2, mode, e3=None):
        agg = torch.mean if self.inter.startswith("mean") else _amin
        parts = [e1, e2] if e3 is None else [e1, e2, e3]
        if self.inter.endswith("-simple"):      # decoders.py:311-319
            return agg(torch.stack(parts), dim=0)
        pre, post = self.pre[mode], self.post[mode]
        hidden = [F.relu(pre.mm(e)) for e in parts]        # decoders.py:289-292
        return post.mm(agg(torch.stack(hidden), dim=0))    # decoders.py:293-299

    # ---- a13: QueryEncoderDecoder.forward ----------------------------------
    def forward(self, formula, queries, source_nodes):
        qt = formula.query_type
        anchors = lambda k: [q.anchor_nodes[k] for q in queries]
        if qt in CHAIN_TYPES:                               # model.py:71-76
            return self.path_score(
                self.encode(source_nodes, formula.target_mode),
                self.encode(anchors(0), formula.anchor_modes[0]),
                formula.rels)
        if qt in FLAT_INTER_TYPES or qt == "3-inter_chain":  # model.py:77-98
            target = self.encode(source_nodes, formula.target_mode)
            e1 = self.project(self.encode(anchors(0), formula.anchor_modes[0]),
                              reverse_relation(formula.rels[0]))
            e2 = self.encode(anchors(1), formula.anchor_modes[1])
            if len(formula.rels[1]) == 2:                   # nested pair of triples
                for r in formula.rels[1][::-1]:
                    e2 = self.project(e2, reverse_relation(r))
            else:
                e2 = self.project(e2, reverse_relation(formula.rels[1]))
            e3 = None
            if qt == "3-inter":
                e3 = self.project(self.encode(anchors(2), formula.anchor_modes[2]),
                                  reverse_relation(formula.rels[2]))
            q = self.intersect(e1, e2, formula.target_mode, e3)
            return F.cosine_similarity(target, q, dim=0, eps=1e-8)
        if qt == "3-chain_inter":                           # model.py:99-109
            target = self.encode(source_nodes, formula.target_mode)
            e1 = self.project(self.encode(anchors(0), formula.anchor_modes[0]),
                              reverse_relation(formula.rels[1][0]))
            e2 = self.project(self.encode(anchors(1), formula.anchor_modes[1]),
                              reverse_relation(formula.rels[1][1]))
            q = self.intersect(e1, e2, formula.rels[0][-1])
            q = self.project(q, reverse_relation(formula.rels[0]))
            return F.cosine_similarity(target, q, dim=0, eps=1e-8)
        return None                                         # model.py: no else branch

    # ---- a14: margin_loss ---------------------------------------------------
    def pick_negatives(self, formula, queries, hard_negatives=False):
        """model.py:113-120; consumes the global ``random`` stream."""
        if "inter" not in formula.query_type and hard_negatives:
            raise Exception("Hard negative examples can only be used with intersection queries")
        if hard_negatives:
            return [random.choice(q.hard_neg_samples) for q in queries]
        if formula.query_type == "1-chain":
            return [random.choice(self.full_lists[formula.target_mode]) for _ in queries]
        return [random.choice(q.neg_samples) for q in queries]

    def margin_loss(self, formula, queries, hard_negatives=False, margin=1, neg_nodes=None):
        """model.py:122-126: two forward passes, hinge, mean."""
        if neg_nodes is None:
            neg_nodes = self.pick_negatives(formula, queries, hard_negatives)
        pos = self.forward(formula, queries, [q.target_node for q in queries])
        neg = self.forward(formula, queries, neg_nodes)
        return torch.clamp(margin - (pos - neg), min=0).mean()


# ---- restated eval batching (utils.py:35-91), used by tests of the eval path --
# ---- f4: GraphSAGE-style encoder (the --depth > 0 path) ---------------------------------------
def mean_aggregate(features, to_neighs, rel, keep_prob=0.5, max_keep=10):
    """MeanAggregator.forward (netquery/aggregators.py:33-68): sample
    ``min(ceil(len * keep_prob), max_keep)`` neighbours per node with ``random.sample`` (the global
    ``random`` stream), build the row-normalised [batch, unique neighbours] mask and multiply it
    with the features of the unique neighbours.  -> [batch, d]"""
    samp_neighs = [set(random.sample(list(to_neigh), min(int(math.ceil(len(to_neigh) * keep_prob)), max_keep)))
                   for to_neigh in to_neighs]
    unique_nodes_list = list(set.union(*samp_neighs))
    unique_nodes = {n: i for i, n in enumerate(unique_nodes_list)}
    mask = torch.zeros(len(samp_neighs), len(unique_nodes))
    column_indices = [unique_nodes[n] for samp_neigh in samp_neighs for n in samp_neigh]
    row_indices = [i for i in range(len(samp_neighs)) for _ in range(len(samp_neighs[i]))]
    mask[row_indices, column_indices] = 1
    mask = mask.div(mask.sum(1, keepdim=True))
    embed_matrix = features(unique_nodes_list, rel[-1])
    if len(embed_matrix.size()) == 1:
        embed_matrix = embed_matrix.unsqueeze(dim=0)
    return mask.mm(embed_matrix)


class OracleSageEncoder(object):
    """Encoder.forward (netquery/encoders.py:103-123) on top of a ``features(nodes, mode) -> [n, d]``
    callable: per outgoing relation type of ``mode`` the mean of sampled neighbours' features, the
    node's own features last, concatenated along the feature axis, compressed by
    ``compress[mode]`` ([d_out, d * (1 + #relations)]) and passed through ReLU.  No L2
    normalisation (unlike DirectEncoder).  -> [d_out, batch]"""

    def __init__(self, features, relations, adj_lists, compress, agg_features=None):
        self.features, self.relations, self.adj_lists = features, relations, adj_lists
        self.agg_features = features if agg_features is None else agg_features   # utils.py:108-119 wires them apart
        self.compress = {m: t.detach().to("cpu", torch.float32).contiguous() for m, t in compress.items()}

    def __call__(self, nodes, mode, keep_prob=0.5, max_keep=10):
        self_feat = self.features(nodes, mode).t()
        neigh_feats = []
        for to_r in self.relations[mode]:
            rel = (mode, to_r[1], to_r[0])
            to_neighs = [[-1] if node == -1 else self.adj_lists[rel][node] for node in nodes]
            to_neighs = [[-1] if len(l) == 0 else l for l in to_neighs]     # null neighbour (encoders.py:112-113)
            neigh_feats.append(mean_aggregate(self.agg_features, to_neighs, rel, keep_prob, max_keep).t())
        neigh_feats.append(self_feat)
        combined = torch.cat(neigh_feats, dim=0)
        return F.relu(self.compress[mode].mm(combined))


def sage_stack(depth, features, relations, adj_lists, compress_layers):
    """get_encoder for depth 1..3 (netquery/utils.py:103-126): layer k's own features and its
    aggregator's features are the transposed, squeezed output of lower layers -- layer 3 reads its
    OWN features from layer 1 and aggregates layer 2 (utils.py:116-119).
    compress_layers: [{mode: [d_out, d * (1 + #relations)]}] per layer."""
    enc1 = OracleSageEncoder(features, relations, adj_lists, compress_layers[0])
    if depth == 1:
        return enc1
    lower1 = lambda nodes, mode: enc1(nodes, mode).t().squeeze()
    enc2 = OracleSageEncoder(lower1, relations, adj_lists, compress_layers[1])
    if depth == 2:
        return enc2
    lower2 = lambda nodes, mode: enc2(nodes, mode).t().squeeze()
    return OracleSageEncoder(lower1, relations, adj_lists, compress_layers[2], agg_features=lower2)


def eval_pairs(formula_queries, offset, batch_size, hard_negatives, one_negative):
    """Build the (queries, targets, lengths) a reference eval batch scores.

    utils.py:48-60 (one sampled negative) / utils.py:78-88 (all negatives): the
    query list is the batch followed by each query repeated once per negative;
    targets are the positives followed by the flattened negatives.
    """
    hi = min(offset + batch_size, len(formula_queries))
    batch = formula_queries[offset:hi]
    pick = (lambda q: q.hard_neg_samples) if hard_negatives else (lambda q: q.neg_samples)
    if one_negative:
        lengths = [1] * len(batch)
        negatives = [random.choice(pick(q)) for q in batch]
    else:
        lengths = [len(pick(q)) for q in batch]
        negatives = [n for q in batch for n in pick(q)]
    rep = [q for i, q in enumerate(batch) for _ in range(lengths[i])]
    return batch + rep, [q.target_node for q in batch] + negatives, lengths


def get_perc_scores(scores, lengths):
    """utils.py:26-33: percentile of each positive among its own negatives."""
    from scipy import stats
    out, cum = [], 0
    neg_scores = scores[len(lengths):]
    for i, length in enumerate(lengths):
        out.append(stats.percentileofscore(neg_scores[cum:cum + length], scores[i]))
        cum += length
    return out


def eval_auc_queries(test_queries, enc_dec, batch_size=1000, hard_negatives=False, seed=0):
    """utils.py:35-68, through ``enc_dec.forward`` with the repeated query list."""
    import numpy as np
    from sklearn.metrics import roc_auc_score
    predictions, labels, formula_aucs = [], [], {}
    random.seed(seed)
    for formula in test_queries:
        f_labels, f_pred = [], []
        formula_queries = test_queries[formula]
        offset = 0
        while offset < len(formula_queries):
            qs, targets, lengths = eval_pairs(formula_queries, offset, batch_size, hard_negatives, True)
            offset += batch_size
            f_labels.extend([1] * len(lengths) + [0] * len(lengths))
            f_pred.extend(enc_dec.forward(formula, qs, targets).data.tolist())
        formula_aucs[formula] = roc_auc_score(f_labels, np.nan_to_num(f_pred))
        labels.extend(f_labels)
        predictions.extend(f_pred)
    return roc_auc_score(labels, np.nan_to_num(predictions)), formula_aucs


def eval_perc_queries(test_queries, enc_dec, batch_size=1000, hard_negatives=False):
    """utils.py:70-91."""
    import numpy as np
    perc = []
    for formula in test_queries:
        formula_queries = test_queries[formula]
        offset = 0
        while offset < len(formula_queries):
            qs, targets, lengths = eval_pairs(formula_queries, offset, batch_size, hard_negatives, False)
            offset += batch_size
            perc.extend(get_perc_scores(enc_dec.forward(formula, qs, targets).data.tolist(), lengths))
    return np.mean(perc)
