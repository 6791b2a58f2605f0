"""Evaluation callers of the scorer: ROC-AUC and percentile rank.

Drop-ins for reference ``netquery/utils.py:26-91`` (``eval_auc_queries``,
``eval_perc_queries``, ``_get_perc_scores``): same arguments, same negative
draws from the global ``random`` stream (``random.seed(seed)`` then one
``random.choice`` per query, utils.py:39,50,53), same metric definitions
(``sklearn.metrics.roc_auc_score`` over ``nan_to_num`` predictions, utils.py:63,66;
``scipy.stats.percentileofscore`` of the positive among the query's negatives,
utils.py:31) -- but each query is lowered ONCE: the reference physically
repeats the query object once per negative (utils.py:58-60,86-88) so that every
anchor is re-gathered and re-projected K times; here a batch is one flat
``QueryBatch`` (ragged target lists) scored by one fused launch plus the
HBM-bound pair-scoring kernel.
"""
import random

import numpy as np

from .query import QueryBatch


def _batch_arrays(formula, queries, negatives, lengths):
    """Flat batch: targets of query i = [positive, its negatives...]."""
    n = len(queries)
    anchors = np.empty((len(formula.anchor_modes), n), dtype=np.int64)
    for k in range(anchors.shape[0]):
        anchors[k] = np.fromiter((q.anchor_nodes[k] for q in queries), dtype=np.int64, count=n)
    lengths = np.asarray(lengths, dtype=np.int64)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths + 1, out=offsets[1:])
    targets = np.empty(int(offsets[-1]), dtype=np.int64)
    targets[offsets[:-1]] = np.fromiter((q.target_node for q in queries), dtype=np.int64, count=n)
    neg_pos = np.ones(len(targets), dtype=bool)
    neg_pos[offsets[:-1]] = False
    targets[neg_pos] = np.fromiter(negatives, dtype=np.int64, count=int(lengths.sum()))
    regular = n > 0 and bool((lengths == lengths[0]).all())
    return QueryBatch(formula, anchors, targets, None if regular else offsets), offsets


def _scores(enc_dec, formula, queries, negatives, lengths):
    """-> (positive scores [n], negative scores flat, offsets into the negatives)."""
    batch, offsets = _batch_arrays(formula, queries, negatives, lengths)
    flat = enc_dec.score_batch(batch).detach().cpu().numpy()
    is_pos = np.zeros(len(flat), dtype=bool)
    is_pos[offsets[:-1]] = True
    neg_offsets = offsets - np.arange(len(offsets))
    return flat[is_pos], flat[~is_pos], neg_offsets


def percentile_of_score(negatives, score):
    """``scipy.stats.percentileofscore(negatives, score)`` (kind='rank'), utils.py:31."""
    from scipy import stats
    return stats.percentileofscore(negatives, score)


def eval_auc_queries(test_queries, enc_dec, batch_size=1000, hard_negatives=False, seed=0):
    """utils.py:35-68 -> (overall_auc, {formula: auc})."""
    from sklearn.metrics import roc_auc_score
    predictions, labels, formula_aucs = [], [], {}
    random.seed(seed)
    for formula in test_queries:
        formula_labels, formula_predictions = [], []
        formula_queries = test_queries[formula]
        offset = 0
        while offset < len(formula_queries):
            batch_queries = formula_queries[offset:offset + batch_size]
            pool = (lambda q: q.hard_neg_samples) if hard_negatives else (lambda q: q.neg_samples)
            negatives = [random.choice(pool(q)) for q in batch_queries]
            offset += batch_size
            pos, neg, _ = _scores(enc_dec, formula, batch_queries, negatives, [1] * len(batch_queries))
            # the reference's order: the batch's positives, then its negatives (utils.py:56-61)
            formula_labels.extend([1] * len(pos) + [0] * len(neg))
            formula_predictions.extend(pos.tolist() + neg.tolist())
        formula_aucs[formula] = roc_auc_score(formula_labels, np.nan_to_num(formula_predictions))
        labels.extend(formula_labels)
        predictions.extend(formula_predictions)
    overall_auc = roc_auc_score(labels, np.nan_to_num(predictions))
    return overall_auc, formula_aucs


def eval_perc_queries(test_queries, enc_dec, batch_size=1000, hard_negatives=False):
    """utils.py:70-91 -> mean percentile rank of the positive among ALL stored negatives."""
    perc_scores = []
    for formula in test_queries:
        formula_queries = test_queries[formula]
        offset = 0
        while offset < len(formula_queries):
            batch_queries = formula_queries[offset:offset + batch_size]
            pool = (lambda q: q.hard_neg_samples) if hard_negatives else (lambda q: q.neg_samples)
            lengths = [len(pool(q)) for q in batch_queries]
            negatives = [n for q in batch_queries for n in pool(q)]
            offset += batch_size
            pos, neg, off = _scores(enc_dec, formula, batch_queries, negatives, lengths)
            for i in range(len(batch_queries)):
                perc_scores.append(percentile_of_score(neg[off[i]:off[i + 1]], pos[i]))
    return np.mean(perc_scores)
