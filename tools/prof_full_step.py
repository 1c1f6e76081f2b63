"""Short program for ncu captures of the full-graph (C3) training step on ONE GPU: builds the sharded trainer at
world 1 and runs 3 steps (the third is the one to profile).  LGCN_BENCH_SHAPE picks the graph."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from lgcn_b200 import sharded  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402

dev = torch.device("cuda:0")
shape = os.environ.get("LGCN_BENCH_SHAPE", "ml25m")
g = synthetic.make_graph(shape, seed=0)
k = synthetic.SHAPES[shape][3]
ops = sharded.CudaOps(g.edges("train"), g.num_users, g.num_items, k, device=dev)
u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
t = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
torch.manual_seed(0)
for _ in range(int(os.environ.get("LGCN_PROF_STEPS", "3"))):
    loss = t.step_sampled(g.num_items, use_graph=False)
torch.cuda.synchronize()
print("ok", float(loss))
