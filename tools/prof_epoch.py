"""Short program for the ncu capture of the persistent step kernel (profiles/): three C2 epochs over device-resident
cluster batches (ML-25M shape, 100 METIS parts) -> three epoch_kernel launches + three flushes.
Usage: python tools/prof_epoch.py   (exits 0 without ncu first, then under ncu)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from bench import NUM_PARTS, load_partition  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402
from lgcn_b200.data.dataset_handler import ClusterData, ClusterLoader, Data  # noqa: E402
from lgcn_b200.models.light_gcn import LightGCN  # noqa: E402
from lgcn_b200.utils import train_test as tt  # noqa: E402

dev = torch.device("cuda:0")
g = synthetic.make_graph("ml25m", seed=0)
train = g.edges("train")
cluster = load_partition(train, g.num_nodes, "ml25m")
cd = ClusterData(Data(edge_index=train.to(dev), num_nodes=g.num_nodes), NUM_PARTS, cluster=cluster)
model = LightGCN(g.num_users, g.num_items, num_layers=3).to(dev)
opt = tt.FusedAdam(model)
loader = ClusterLoader([d for d in cd.parts], shuffle=False)
torch.manual_seed(0)
for _ in range(3):
    loss = tt.train(model, opt, loader, dev)
torch.cuda.synchronize()
print("ok", loss)
