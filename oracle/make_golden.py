"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz with the REAL reference.

Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

For every configuration below a seeded case (oracle/cases.py) is evaluated by
the unmodified reference modules (oracle/ref_shim.py) and the inputs, the
reference outputs and the reference's index trace are frozen:

  exp/<structure>/pos      forward(formula, queries, positives)       model.py:70
  exp/<structure>/neg      forward(formula, queries, first negatives)
  exp/<structure>/eval     forward(formula, batch + repeats, positives + all negatives)
                           -- the eval_perc_queries call shape, utils.py:86-88
  exp/<structure>/loss     margin_loss(formula, queries) after random.seed(LOSS_SEED)
                           (model.py:112-127; consumes the global random stream)
  exp/<structure>/hard     margin_loss(..., hard_negatives=True), inter structures only
  meta exp.trace/<structure>  the order of (mode, rows) lookups and relation-key reads
                           during the `pos` call: the "indices bit-exact" contract.
"""
import os
import random
import sys

import numpy as np
import torch

from graphqembed_b200.synth import STRUCTURES

from . import ref_shim
from .cases import make_case

LOSS_SEED = 20260917
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> make_case kwargs
CONFIGS = {
    "d32_bilinear_mean": dict(seed=11, d=32, decoder="bilinear", inter="mean"),
    "d32_bilinear_min": dict(seed=12, d=32, decoder="bilinear", inter="min"),
    "d32_bilinear_mean-simple": dict(seed=13, d=32, decoder="bilinear", inter="mean-simple"),
    "d32_bilinear_min-simple": dict(seed=14, d=32, decoder="bilinear", inter="min-simple"),
    "d32_transe_mean": dict(seed=15, d=32, decoder="transe", inter="mean"),
    "d32_transe_min-simple": dict(seed=16, d=32, decoder="transe", inter="min-simple"),
    "d32_bilinear-diag_min": dict(seed=17, d=32, decoder="bilinear-diag", inter="min"),
    "d32_bilinear-diag_mean-simple": dict(seed=18, d=32, decoder="bilinear-diag", inter="mean-simple"),
    "d64_bilinear_min": dict(seed=19, d=64, decoder="bilinear", inter="min", n_queries=70, n_neg=3),
    "d128_bilinear_mean": dict(seed=20, d=128, decoder="bilinear", inter="mean", n_modes=2, n_rel_pairs=2,
                               n_queries=16, n_neg=2),
    # several 128-row tiles of the d = 256 tensor-core kernel that produces the headline number
    # (chains: 2 x 400 rows = 7 tiles; intersections: 4 tiles, the last one ragged)
    "d256_bilinear_mean": dict(seed=21, d=256, decoder="bilinear", inter="mean", n_modes=2, n_rel_pairs=2,
                               n_queries=400, n_neg=2),
}


def _normalise_trace(trace, decoder):
    """JSON-able copy of the reference's access log.  The vector decoders read
    ``self.vecs[rel]`` twice per use (once for the value, once for ``.size(0)``:
    decoders.py:203,208,231,236), so each such pair of reads is ONE logical use."""
    out, i = [], 0
    while i < len(trace):
        t = trace[i]
        if t[0] == "rows":
            out.append(list(t))
        else:
            out.append(["rel", list(t[1])])
            if decoder != "bilinear":
                assert trace[i + 1] == t, "expected the reference's double read of a relation vector"
                i += 1
        i += 1
    return out


def reference_outputs(case):
    trace = []
    model, g = ref_shim.build_reference_model(case.tables, case.kg.node_maps(), case.kg.relations, case.rel_params,
                                              case.decoder, case.inter, case.pre, case.post,
                                              full_lists=case.kg.full_lists(), trace=trace)
    exp, traces = {}, {}
    with torch.no_grad():
        for s in case.batches:
            formula = case.formula(s, cls=g.Formula)
            queries = case.queries(s, cls=g.Query)
            b = case.batches[s]
            del trace[:]
            pos = model.forward(formula, queries, [q.target_node for q in queries])
            traces[s] = _normalise_trace(trace, case.decoder)
            neg = model.forward(formula, queries, [int(x) for x in b["negs"][:, 0]])
            lengths = [len(q.neg_samples) for q in queries]
            rep = [q for i, q in enumerate(queries) for _ in range(lengths[i])]
            ev = model.forward(formula, queries + rep,
                               [q.target_node for q in queries] + [n for q in queries for n in q.neg_samples])
            random.seed(LOSS_SEED)
            loss = model.margin_loss(formula, queries)
            exp[s + "/pos"] = pos.numpy().copy()
            exp[s + "/neg"] = neg.numpy().copy()
            exp[s + "/eval"] = ev.numpy().copy()
            exp[s + "/loss"] = np.array(loss.item(), dtype=np.float32)
            if "inter" in s:
                random.seed(LOSS_SEED)
                hard = model.margin_loss(formula, queries, hard_negatives=True)
                exp[s + "/hard"] = np.array(hard.item(), dtype=np.float32)
    exp["trace"] = traces
    return exp


def main(argv=None):
    if not ref_shim.available():
        sys.exit("reference tree not mounted; golden vectors can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)     # one fixed reduction order for the frozen vectors
    only = set((argv if argv is not None else sys.argv[1:]))
    for name, kw in CONFIGS.items():
        if only and name not in only:
            continue
        case = make_case(**kw)
        exp = reference_outputs(case)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        case.save(path, exp)
        print("%-34s %7.1f KiB" % (name, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
