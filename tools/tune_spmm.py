"""Tuning aid: time full-graph K-layer propagation fwd / bwd and a full training step for the library
selected by LGCN_LIB_PATH (variants built by lgcn_b200.build.build_variant)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from lgcn_b200 import _lib  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402
from lgcn_b200.models.light_gcn import LightGCN  # noqa: E402
from lgcn_b200.utils import train_test as tt  # noqa: E402

dev = torch.device("cuda:0")
g = synthetic.make_graph("ml25m", seed=0)
tr = g.edges("train").to(dev)
model = LightGCN(g.num_users, g.num_items, num_layers=3).to(dev)
opt = tt.FusedAdam(model)
G = model.graph(tr)
L = _lib.lib()
s = _lib.stream_ptr(dev)
b = opt.buffers
uw, iw = model.user_embedding.weight, model.item_embedding.weight
neg = torch.randint(0, g.num_items, (G.num_triplets,), device=dev)
b.ensure_triplets(G.num_triplets)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); z.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(z))
    return float(np.median(ts))


fwd = timed(lambda: L.lgcn_propagate_fwd(G.ref, uw.data_ptr(), iw.data_ptr(), 3, b.final_emb.data_ptr(), b.rnorm.data_ptr(),
                                         b.work.data_ptr(), b.work.numel() * 4, s))
bpr = timed(lambda: L.lgcn_bpr_fwd_bwd(G.ref, b.final_emb.data_ptr(), b.rnorm.data_ptr(), neg.data_ptr(), b.grad_final.data_ptr(),
                                       b.neg_count.data_ptr(), b.trip_scratch.data_ptr(), b.accum.data_ptr(), s))
bwd = timed(lambda: L.lgcn_propagate_bwd(G.ref, b.grad_final.data_ptr(), 3, uw.data_ptr(), iw.data_ptr(), b.neg_count.data_ptr(),
                                         1e-9, b.grad_e0.data_ptr(), b.accum.data_ptr(), b.work.data_ptr(), b.work.numel() * 4, s))
step = timed(lambda: tt.train_step(model, opt, tr, neg), iters=10)
print(os.path.basename(_lib.LIB_PATH), f"fwd3 {fwd:.3f} ms  bpr {bpr:.3f} ms  bwd3 {bwd:.3f} ms  step {step:.3f} ms")
