"""ncu driver for the scoring kernels: one wave of CTAs (148 x 128 users) against all ML-25M items."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from lgcn_b200.data import synthetic  # noqa: E402
from lgcn_b200.utils import recommend as rec  # noqa: E402

dev = torch.device("cuda:0")
U, I = 148 * 128, 59047
u0, i0 = synthetic.init_embeddings(U, I, 64, 0)
ue, ie = u0.to(dev), i0.to(dev)
algo = rec.SCORE_TENSOR if (len(sys.argv) < 2 or sys.argv[1] == "tc") else rec.SCORE_FFMA
for _ in range(3):
    ids, vals = rec.score_topk(ue, ie, 20, True, algo=algo)
torch.cuda.synchronize()
a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); rec.score_topk(ue, ie, 20, True, algo=algo); z.record(); torch.cuda.synchronize()
print("ok", a.elapsed_time(z), "ms for one wave")
