"""Host-time profile of QueryEncoderDecoder.margin_loss(formula, StoreSlice) at the reference's
training batch (512 queries): where the microseconds of one call go (run on the GPU box)."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import graphqembed_b200 as gqe  # noqa: E402
from graphqembed_b200.store import QueryStore  # noqa: E402
from graphqembed_b200.synth import SynthKG  # noqa: E402
from helpers import build_package_model  # noqa: E402
from oracle.cases import make_case  # noqa: E402

case = make_case(seed=1, d=128, decoder="bilinear", inter="mean", n_queries=2048, n_neg=4, nodes_per_mode=5000)
model = build_package_model(case)
raw = []
for s in case.batches:
    b = case.batches[s]
    for i in range(len(b["target"])):
        negs = [int(x) for x in b["negs"][i]]
        raw.append((SynthKG.query_graph(s, b["rels"], b["target"][i], b["anchors"][:, i]), negs, negs if "inter" in s else None))
store = QueryStore.from_records(raw)
model.negative_rng = np.random.default_rng(0)
for s in ("1-chain", "3-inter"):
    f = case.formula(s, cls=gqe.Formula)
    sl = store[f].window(0, 512)
    for check in (True, False):
        model.check_indices = check
        with torch.no_grad():
            for _ in range(20):
                model.margin_loss(f, sl)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 300
            for _ in range(n):
                loss = model.margin_loss(f, sl)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        print("%s check_indices=%s: %.1f us/call host, %.1f us/call incl. final sync" % (s, check, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
    pr = cProfile.Profile()
    with torch.no_grad():
        pr.enable()
        for _ in range(200):
            model.margin_loss(f, sl)
        pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
