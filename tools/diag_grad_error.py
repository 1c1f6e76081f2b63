"""Where does the fp32 error of the full-graph step sit?  One GPU, LGCN_BENCH_SHAPE graph: the step's dL/dfinal (G)
and dL/dE0 against the float64 restatement (tests/fp64_ref.py), error by row degree."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import lgcn_b200  # noqa: E402,F401
import fp64_ref  # noqa: E402
from lgcn_b200 import sharded  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402

dev = torch.device("cuda:0")
shape = os.environ.get("LGCN_BENCH_SHAPE", "ml25m_x10")
nu, ni, _, k = synthetic.SHAPES[shape]
g = synthetic.make_graph(shape, seed=0)
train = g.edges("train")
del g
ops = sharded.CudaOps(train, nu, ni, k, device=dev)
u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
t = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
gen = torch.Generator().manual_seed(29)
neg = torch.randint(0, ni, (ops.num_triplets,), generator=gen).to(dev)
loss = float(t.step(neg))
e0 = torch.cat([u0, i0]).to(dev).double()
tr = train.to(dev)
rl, rgrad, rfinal = fp64_ref.step_loss_and_grad(tr, e0, k, nu, neg)
# G in fp64: grad = (sum_j (A^T)^j G)/(K+1)^2 + reg  -> recompute G directly
row, col, dis = fp64_ref.norm_weights(tr, nu + ni)
deg = torch.bincount(col, minlength=nu + ni) + torch.bincount(row, minlength=nu + ni)
print("loss", loss, "fp64", rl)
for name, got, want in (("final^", ops.final, rfinal / rfinal.norm(dim=1, keepdim=True)), ("grad", ops.grad, rgrad)):
    d = (got.double() - want).abs()
    den = float(want.abs().max())
    rowmax = d.max(dim=1).values
    top = torch.topk(rowmax, 5)
    print(f"{name}: normwise {float(d.max()) / den:.3e}; max|ref| {den:.3e}; worst rows {top.indices.tolist()} "
          f"deg {deg[top.indices].tolist()} err {[f'{x:.2e}' for x in top.values.tolist()]} "
          f"ref-row-max {[f'{float(want[i].abs().max()):.2e}' for i in top.indices.tolist()]}")
    for lo, hi in ((0, 100), (100, 10_000), (10_000, 10**9)):
        m = (deg >= lo) & (deg < hi)
        if m.any():
            print(f"   deg [{lo},{hi}): rows {int(m.sum())} max err {float(rowmax[m].max()):.3e}  max|ref| {float(want[m].abs().max()):.3e}")
