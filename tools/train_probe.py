"""Native training step at the reference's shape (1-chain, Bilinear, d=128, batch 512): a few steps for an
ncu launch list (run under gpurun:  ncu --metrics gpu__time_duration.sum --csv ... python tools/train_probe.py)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphqembed_b200 as gqe  # noqa: E402
from graphqembed_b200 import _lib  # noqa: E402
from graphqembed_b200.synth import bio_shaped  # noqa: E402

structure = sys.argv[1] if len(sys.argv) > 1 else "1-chain"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
device = torch.device("cuda", 0)
kg = bio_shaped(seed=0)
rng = np.random.RandomState(7)
rels = kg.sample_rels(structure, rng)
b = kg.sample_batch(structure, rels, batch, 1, rng)
torch.manual_seed(0)
tables = [torch.randn(kg.sizes[m] + 2, d, device=device) / d for m in kg.modes]
mats = [torch.randn(d, d, device=device) * 0.06 for _ in kg.rel_keys]
pre = [torch.randn(d, d, device=device) * 0.06 for _ in kg.modes]
post = [torch.randn(d, d, device=device) * 0.06 for _ in kg.modes]
ctx = gqe.Context(0, torch.cuda.current_stream().cuda_stream)
ctx.bind_tables([x.data_ptr() for x in tables], [x.size(0) for x in tables], d)
ctx.bind_relations(_lib.DECODER_ID["bilinear"], [r.data_ptr() for r in mats], d)
ctx.bind_intersection(_lib.INTER_ID["mean"], [x.data_ptr() for x in pre], [x.data_ptr() for x in post], d)
lookup = gqe.RowLookup(kg.node_ids)
maps = lookup.device_maps(kg.modes, [x.size(0) for x in tables], device)
ctx.bind_node_maps(maps[0], maps[1], maps[2])
plan = gqe.lower_formula(gqe.Formula(structure, rels), {m: i for i, m in enumerate(kg.modes)},
                         {r: i for i, r in enumerate(kg.rel_keys)})
anchors = torch.from_numpy(np.ascontiguousarray(b["anchors"], dtype=np.int32)).to(device)
pairs = torch.from_numpy(np.stack([b["target"], b["negs"][:, 0]], 1).astype(np.int32)).to(device)
loss = torch.zeros(1, device=device)
hyper = _lib.AdamHyper(lr=0.01)


def step():
    ctx.train_step_device(plan, batch, anchors.data_ptr(), pairs.data_ptr(), 1.0, hyper, loss.data_ptr(), nodes=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
l0 = ctx.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    step()
e1.record()
t_issue = (time.perf_counter() - t0) * 1e3 / steps
torch.cuda.synchronize()
print("%s batch %d d %d: %.4f ms/step on the device, %.4f ms/step host issue time, %.1f kernels/step, loss %.5f"
      % (structure, batch, d, e0.elapsed_time(e1) / steps, t_issue, (ctx.launch_count() - l0) / steps, loss.item()))
