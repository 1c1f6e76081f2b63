"""Short program for ncu captures (profiles/): on the ML-25M-shaped graph run
  (a) 2 full-graph training steps   (K1 x3, K3 pass A/B, K2 x3, K6 on 22.5 M edges), then
  (b) 2 training steps on the largest and 2 on a median Cluster-GCN batch.
Usage: python tools/prof_step.py   (exits 0 without ncu first, then under ncu)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from bench import NUM_PARTS, load_partition  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402
from lgcn_b200.data.dataset_handler import ClusterData, Data  # noqa: E402
from lgcn_b200.models.light_gcn import LightGCN  # noqa: E402
from lgcn_b200.utils import train_test as tt  # noqa: E402

dev = torch.device("cuda:0")
g = synthetic.make_graph("ml25m", seed=0)
train = g.edges("train")
cluster = load_partition(train, g.num_nodes, "ml25m")
tr = train.to(dev)
cd = ClusterData(Data(edge_index=tr, num_nodes=g.num_nodes), NUM_PARTS, cluster=cluster)
sizes = np.array([d.edge_index.shape[1] for d in cd.parts])
big, med = cd.parts[int(sizes.argmax())], cd.parts[int(np.argsort(sizes)[len(sizes) // 2])]
model = LightGCN(g.num_users, g.num_items, num_layers=3).to(dev)
opt = tt.FusedAdam(model)
torch.manual_seed(0)
for ei in (tr, tr, big.edge_index, big.edge_index, med.edge_index, med.edge_index):
    loss = tt.train_step(model, opt, ei)
torch.cuda.synchronize()
print("ok", float(loss), "full E", tr.shape[1], "big E", int(sizes.max()), "median E", med.edge_index.shape[1])
