"""Parity at BASELINE.json's FULL sizes (ML-25M shape: 162,541 users x 59,047 movies, 22.5 M train edges).

  C2  one Cluster-GCN epoch -- the committed 100-part METIS vector, INCLUDING the hub cluster (606 k edges,
      10 k split slots) -- through train() / the persistent step kernel against the CPU oracle's TrainState
      on the same negatives: all 100 batch losses and the final weights (utils/train_test.py:66-103).
  C3  one full-graph training step (owner-computes BPR, sharded orchestration at world 1) against the CPU
      oracle on a 1/8 edge sample of the same node set, and at full size against fp64 torch on the device.
  C4  full-rank scoring of every user; 512 sampled users against an fp64 brute force
      (utils/recommend.py:39-61 batched, utils/train_test.py:191-197).
"""
import os

import numpy as np
import pytest
import torch

import lgcn_b200  # noqa: F401
from conftest import ADAM_STEP_ATOL, max_abs, normwise
from lgcn_b200.data import synthetic
from lgcn_b200.data.dataset_handler import ClusterData, ClusterLoader, Data
from lgcn_b200.models.light_gcn import LightGCN
from lgcn_b200.utils import recommend as rec
from lgcn_b200.utils import train_test as tt
from oracle import reference_path as ref

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ml25m():
    g = synthetic.make_graph("ml25m", seed=0)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    return g, g.edges("train"), u0, i0


def _model(g, k, u0, i0):
    m = LightGCN(g.num_users, g.num_items, num_layers=k).to(DEV)
    with torch.no_grad():
        m.user_embedding.weight.copy_(u0)
        m.item_embedding.weight.copy_(i0)
    return m


def test_c2_full_epoch_ml25m_metis100_incl_hub_cluster_vs_oracle(ml25m):
    g, train, u0, i0 = ml25m
    z = np.load(os.path.join(REPO, "tests", "golden", "ml25m_seed0_metis100.npz"))
    assert int(z["train_checksum"]) == int((train[0] * 31 + train[1]).sum()), "fixture belongs to another graph"
    cluster = torch.from_numpy(z["cluster"].astype(np.int64))
    n, k = g.num_nodes, 3
    cd = ClusterData(Data(edge_index=train.to(DEV), num_nodes=n), 100, cluster=cluster)
    parts = [d for d in cd.parts if int((d.edge_index[0] < g.num_users).sum()) > 0]
    sizes = [int(d.edge_index.shape[1]) for d in parts]
    assert len(parts) >= 99 and max(sizes) > 500_000, "the hub cluster must be part of the epoch"
    m = _model(g, k, u0, i0)
    opt = tt.FusedAdam(m)
    loader = ClusterLoader(parts, shuffle=False)
    torch.manual_seed(11)
    epoch_loss = tt.train(m, opt, loader, DEV)                    # ONE persistent launch (lgcn_train_steps_sparse)
    assert int(opt.step_count) == len(parts) and not opt.pending
    losses = opt.losses[: len(parts)].cpu().double()
    # the negatives train() drew: one randint over the run, in loader order (utils/helpers.py:79-80)
    trip = [int((d.edge_index[0] < g.num_users).sum()) for d in parts]
    torch.manual_seed(11)
    neg_all = torch.randint(0, g.num_items, (sum(trip),), device=DEV).cpu()
    negs = torch.split(neg_all, trip)
    st = ref.TrainState(u0, i0, k)
    want = [st.step(d.edge_index.cpu(), ng) for d, ng in zip(parts, negs)]
    rel = [abs(float(a) - w) / abs(w) for a, w in zip(losses, want)]
    hub = int(np.argmax(sizes))
    assert max(rel) < 1e-4, f"worst batch {int(np.argmax(rel))} rel {max(rel):.2e} (hub cluster is batch {hub}: {rel[hub]:.2e})"
    w = torch.tensor(sizes, dtype=torch.float64)
    assert abs(epoch_loss - float((torch.tensor(want, dtype=torch.float64) * w).sum() / w.sum())) < 1e-5 * abs(epoch_loss)
    steps = len(parts)
    du = (m.user_embedding.weight.detach().cpu() - st.user_w.detach()).abs()
    di = (m.item_embedding.weight.detach().cpu() - st.item_w.detach()).abs()
    moved = float((st.item_w.detach() - i0).abs().max())
    print(f"C2 epoch: max rel loss err {max(rel):.2e}; weights max|d| user {float(du.max()):.2e} item {float(di.max()):.2e} "
          f"(largest oracle movement {moved:.2e}); 99.99th pct {float(torch.quantile(di.flatten()[::7].double(), 0.9999)):.2e}")
    assert float(du.max()) < steps * ADAM_STEP_ATOL and float(di.max()) < steps * ADAM_STEP_ATOL
    # the bulk of the table is far inside the Adam-noise bound
    assert float((du > 1e-5).double().mean()) < 1e-3 and float((di > 1e-5).double().mean()) < 1e-3


def test_c3_full_graph_step_ml25m_vs_fp64_on_device(ml25m):
    """One full-graph training step at full size: forward, loss and dL/dE0 against float64 torch ops on the device
    (the reference's op sequence: gcn_norm + index_select / mul / index_add_, cosine BPR), same negatives."""
    from lgcn_b200 import sharded
    g, train, u0, i0 = ml25m
    k, n, nu = 3, g.num_nodes, g.num_users
    tr = train.to(DEV)
    ops = sharded.CudaOps(tr, g.num_users, g.num_items, k)
    uw, iw = u0.to(DEV).clone(), i0.to(DEV).clone()
    trainer = sharded.ShardedTrainer(ops, uw, iw, sharded.Comm())
    gen = torch.Generator().manual_seed(3)
    p = int((train[0] < nu).sum())
    neg = torch.randint(0, g.num_items, (p,), generator=gen).to(DEV)
    plain_final = trainer.propagate_only().clone()                 # inference path: the plain final rows
    loss = float(trainer.step(neg).item())
    grad = ops.grad.clone()
    # fp64 reference on the device
    row, col = tr[0], tr[1]
    deg = torch.bincount(col, minlength=n).double()
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    w = (dis[row] * dis[col])[:, None]
    e0 = torch.cat([u0, i0]).to(DEV).double().requires_grad_(True)
    x, acc = e0, e0
    for _ in range(k):
        x = torch.zeros(n, 64, device=DEV, dtype=torch.float64).index_add_(0, col, w * x.index_select(0, row))
        acc = acc + x
    final = acc / float((k + 1) ** 2)
    m = row < nu
    u, pos, ng = row[m], col[m], neg + nu

    def nrm(t):
        return t / t.norm(dim=1, keepdim=True)
    total = 0.0
    reg = 0.0
    chunk = 1 << 21                                              # keep the [P,64] fp64 temporaries bounded
    for b in range(0, p, chunk):
        s = slice(b, min(p, b + chunk))
        uf, pf, nf = nrm(final[u[s]]), nrm(final[pos[s]]), nrm(final[ng[s]])
        x10 = 10.0 * ((uf * pf).sum(1) - (uf * nf).sum(1))
        total = total + torch.nn.functional.softplus(x10).sum()
        reg = reg + (e0[u[s]] ** 2).sum() + (e0[pos[s]] ** 2).sum() + (e0[ng[s]] ** 2).sum()
    want = -total / (10.0 * p) + 5e-3 * reg / (64.0 * p)
    want.backward()
    assert abs(loss - float(want)) < 1e-5 * abs(float(want))
    # the step keeps final / ||final|| (every copy) and 1/||final|| (owned rows)
    fn = final.detach().norm(dim=1, keepdim=True)
    assert normwise(ops.final, final.detach() / fn) < 1e-5
    assert normwise(ops.rnorm, 1.0 / fn.flatten()) < 1e-5
    assert normwise(plain_final, final.detach()) < 1e-5
    assert normwise(grad, e0.grad) < 1e-5
    # Adam on those gradients: first step => every element moves by lr * sign(g) (up to eps effects)
    moved = torch.cat([uw, iw]) - torch.cat([u0, i0]).to(DEV)
    gn = float(e0.grad.norm())
    gc = e0.grad * min(1.0, 1.0 / (gn + 1e-6))
    want_move = -1e-3 * gc / (gc.abs() + 1e-8)
    assert float((moved.double() - want_move).abs().max()) < 2e-5


def test_c4_full_rank_scoring_ml25m_512_sampled_users_vs_fp64(ml25m):
    g, train, u0, i0 = ml25m
    ue, ie = u0.to(DEV), i0.to(DEV)
    tr = train.to(DEV)
    ptr, idx = rec.exclusion_csr(tr, g.num_users)
    k = 20
    ids, vals = rec.score_topk(ue, ie, k, True, ptr, idx)          # all 162,541 users, tcgen05 kernel
    assert ids.shape == (g.num_users, k)
    gen = torch.Generator().manual_seed(9)
    pick = torch.randperm(g.num_users, generator=gen)[:512].to(DEV)
    # the heaviest users (longest exclusion rows) are part of the sample
    deg = ptr[1:] - ptr[:-1]
    pick = torch.unique(torch.cat([pick, torch.topk(deg, 8).indices]))
    un = torch.nn.functional.normalize(ue[pick].double(), dim=1)
    inn = torch.nn.functional.normalize(ie.double(), dim=1)
    sc = un @ inn.t()
    for j, u in enumerate(pick.tolist()):
        sc[j, idx[int(ptr[u]): int(ptr[u + 1])].long()] = -float("inf")
    ov, oi = torch.topk(sc, k, dim=1)
    got_v, got_i = vals[pick].double(), ids[pick].long()
    tol = 1e-5
    assert float((got_v - ov).abs().max()) <= tol
    gap_prev = torch.ones_like(ov, dtype=torch.bool)
    gap_prev[:, 1:] = (ov[:, :-1] - ov[:, 1:]) > 2 * tol
    gap_next = torch.ones_like(gap_prev)
    gap_next[:, :-1] = gap_prev[:, 1:]
    safe = gap_prev & gap_next
    safe[:, -1] = False
    assert torch.equal(got_i[safe], oi[safe]) and float(safe.double().mean()) > 0.8
    # never a masked train item, lists sorted
    for j, u in enumerate(pick.tolist()[:64]):
        assert not torch.isin(got_i[j], idx[int(ptr[u]): int(ptr[u + 1])].long()).any()
    assert (vals[:, :-1] >= vals[:, 1:]).all()
