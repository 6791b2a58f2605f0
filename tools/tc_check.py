"""Debug aid: tensor-core path vs the exact-fp32 kernels on identical inputs."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_package_model, query_batch  # noqa: E402
from oracle.cases import make_case  # noqa: E402

dims = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["128", "256"])]
inters = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mean", "min", "mean-simple", "min-simple"]
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 300
compose = sys.argv[4] if len(sys.argv) > 4 else "auto"
worst = 0.0
for d in dims:
    for inter in inters:
        case = make_case(seed=3 + d, d=d, decoder="bilinear", inter=inter, n_queries=nq, n_neg=3, nodes_per_mode=500)
        model = build_package_model(case)
        model.compose = compose
        for s in case.batches:
            b = case.batches[s]
            targets = np.concatenate([b["target"][:, None], b["negs"]], axis=1)
            qb = query_batch(case, s, targets)
            model.precision = "fp32"
            ref = model.score_batch(qb).cpu().numpy()
            model.precision = "bf16x3"
            t0 = time.time()
            got = model.score_batch(qb).cpu().numpy()
            err = float(np.abs(got - ref).max())
            qb2 = query_batch(case, s, targets[:, :2])
            model.precision = "fp32"
            l_ref = model.margin_loss_batch(qb2).item()
            model.precision = "bf16x3"
            l_got = model.margin_loss_batch(qb2).item()
            worst = max(worst, err, abs(l_got - l_ref))
            print("d=%d %-11s %-14s max|tc-fp32|=%.3e loss %.6f vs %.6f (%.2fs)" % (d, inter, s, err, l_got, l_ref, time.time() - t0),
                  flush=True)
print("WORST", worst)
