"""The STAGE instantiation of the fused tensor-core kernel (node-type-sharded tables: a helper warp fetches
remote operand rows by TMA into a staging area, the workers gather them from there) on ONE GPU: with
GQE_FORCE_STAGE=1 every table is treated as remote, so every operand of every tile but a CTA's first goes
through the helper -- the scores must be the bits of the ordinary kernel.  (Environment knobs are read when the
library is loaded: the forced run is a child process.)  On two GPUs the same kernel runs in
tests/test_gpu_sharded.py and in every sharded leg of bench.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import graphqembed_b200 as gqe
from graphqembed_b200.synth import STRUCTURES
from helpers import build_package_model
from oracle.cases import make_case

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grouped_scores(inter):
    case = make_case(seed=321, d=256, decoder="bilinear", inter=inter, n_queries=3300, n_neg=1, nodes_per_mode=1500)
    model = build_package_model(case)
    batches = []
    for s in STRUCTURES:
        b = case.batches[s]
        f = case.formula(s, cls=gqe.Formula)
        batches.append(gqe.QueryBatch(f, b["anchors"], np.stack([b["target"], b["negs"][:, 0]], 1).reshape(-1)))
    loss, scores = model.margin_loss_grouped(batches, margin=1, return_scores=True)
    return loss.item(), scores.cpu().numpy()


@pytest.mark.parametrize("inter", ["mean", "min-simple"])
def test_forced_staging_is_bit_identical(inter, tmp_path):
    want_loss, want = _grouped_scores(inter)          # 7 formulas x 3300 queries: ~330 tiles, several per CTA
    out = str(tmp_path / "staged.npz")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np; import test_gpu_stage as t; "
            "l, s = t._grouped_scores(%r); np.savez(%r, loss=l, scores=s)" % (ROOT, os.path.join(ROOT, "tests"), inter, out))
    env = dict(os.environ, GQE_FORCE_STAGE="1")
    proc = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=300)
    assert proc.returncode == 0, proc.stdout[-2000:]
    got = np.load(out)
    np.testing.assert_array_equal(got["scores"], want)
    assert float(got["loss"]) == want_loss
