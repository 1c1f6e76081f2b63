"""CPU: host-side logic that needs no device -- generator contract, loader semantics, graph cache
keying, undirected/ exclusion helpers, state_dict layout."""
import numpy as np
import torch

import lgcn_b200  # noqa: F401
from lgcn_b200.data import dataset_handler as dh
from lgcn_b200.data import synthetic
from lgcn_b200.models.light_gcn import GraphCache, LightGCN
from lgcn_b200.utils import recommend as rec
from lgcn_b200.utils.helpers import get_triplets_indices
from oracle import pyg_restated as pyg
from oracle import reference_path as ref


def test_synthetic_graph_contract():
    g = synthetic.make_graph("ml100k", seed=0)
    ei, n = g.edge_index, g.num_nodes
    assert ei.shape == (2, 200_000)
    key = ei[0] * n + ei[1]
    assert (key[1:] > key[:-1]).all()                                 # to_undirected order, no dups
    half = ei.shape[1] // 2
    assert (ei[0, :half] < g.num_users).all() and (ei[1, :half] >= g.num_users).all()
    assert torch.equal(pyg.to_undirected(ei[:, :half]), ei)
    assert torch.equal(dh.to_undirected(ei[:, :half]), ei)
    deg = torch.bincount(ei[1], minlength=n)
    assert int(deg.min()) >= 1                                        # every user and item appears
    allidx = torch.sort(torch.cat([g.train_idx, g.val_idx, g.test_idx]))[0]
    assert torch.equal(allidx, torch.arange(ei.shape[1]))
    assert g.train_idx.numel() == 180_000 and (g.train_idx[1:] > g.train_idx[:-1]).all()
    g2 = synthetic.make_graph("ml100k", seed=0)
    assert torch.equal(g2.edge_index, ei) and torch.equal(g2.train_idx, g.train_idx)
    # directed split => asymmetric train graph
    tr = g.edges("train")
    fwd = set((tr[0] * n + tr[1]).tolist())
    assert any((int(c) * n + int(r)) not in fwd for r, c in zip(tr[0, :200].tolist(), tr[1, :200].tolist()))


def test_triplet_helper_matches_oracle():
    g = synthetic.make_graph("tiny", seed=0)
    tr = g.edges("train")
    torch.manual_seed(3)
    u, p, n = get_triplets_indices(tr, g.num_users, g.num_items, torch.device("cpu"))
    ou, op = ref.triplet_users_pos(tr, g.num_users)
    torch.manual_seed(3)
    on = ref.sample_negative(ou.numel(), g.num_items)
    assert torch.equal(u, ou) and torch.equal(p, op) and torch.equal(n, on)


def test_cluster_loader_semantics():
    parts = [dh.Data(edge_index=torch.full((2, i + 1), i), num_nodes=9) for i in range(10)]
    loader = dh.ClusterLoader(parts, shuffle=True)
    torch.manual_seed(0)
    a = [int(b.edge_index[0, 0]) for b in loader]
    torch.manual_seed(0)
    ref_loader = pyg.DataLoader([pyg.Data(edge_index=p.edge_index) for p in parts], batch_size=1, shuffle=True)
    b = [int(x.edge_index[0, 0]) for x in ref_loader]
    assert a == b and sorted(a) == list(range(10)) and a != list(range(10))
    assert [int(x.edge_index[0, 0]) for x in dh.ClusterLoader(parts, shuffle=False)] == list(range(10))
    d = parts[0].to("cpu")
    assert d is parts[0] and d.num_nodes == 9


def test_graph_cache_keys_on_tensor_identity_and_version(monkeypatch):
    built = []

    class FakeGraph:
        def __init__(self, ei, nu, ni):
            built.append(ei)
            self.num_users, self.num_items = nu, ni

    import lgcn_b200.models.light_gcn as lg
    monkeypatch.setattr(lg, "Graph", FakeGraph)
    c = GraphCache(capacity=2)
    a, b = torch.zeros(2, 3, dtype=torch.int64), torch.ones(2, 3, dtype=torch.int64)
    g1 = c.get(a, 1, 1)
    assert c.get(a, 1, 1) is g1 and len(built) == 1
    a[0, 0] = 5                                                        # in-place edit => rebuild
    assert c.get(a, 1, 1) is not g1 and len(built) == 2
    c.get(b, 1, 1); c.get(torch.zeros(2, 1, dtype=torch.int64), 1, 1)
    assert len(c._d) == 2                                              # LRU capacity


def test_state_dict_layout_is_the_reference_checkpoint_layout(tmp_path):
    m = LightGCN(11, 13, num_layers=3)
    sd = m.state_dict()
    assert list(sd.keys()) == ["user_embedding.weight", "item_embedding.weight"]
    assert sd["user_embedding.weight"].shape == (11, 64) and sd["item_embedding.weight"].shape == (13, 64)
    path = tmp_path / "best_model.pth"
    torch.save({"user_embedding.weight": torch.ones(11, 64), "item_embedding.weight": torch.zeros(13, 64)}, path)
    m.load_state_dict(torch.load(path, map_location="cpu"))            # utils/train_test.py:279-280
    assert float(m.user_embedding.weight.sum()) == 11 * 64
    assert abs(float(LightGCN(2000, 10).user_embedding.weight.std()) - 0.01) < 1e-3
    assert len(m.convs) == 3 and m.num_users == 11 and m.num_items == 13 and m.dim_h == 64


def test_exclusion_csr_on_cpu_tensors():
    g = synthetic.make_graph("tiny", seed=0)
    tr = g.edges("train")
    ptr, idx = rec.exclusion_csr(tr, g.num_users)
    for u in (0, 7, g.num_users - 1):
        want = torch.sort(tr[1, tr[0] == u] - g.num_users)[0]
        assert torch.equal(idx[ptr[u]:ptr[u + 1]].long(), want)
    assert int(ptr[-1]) == int((tr[0] < g.num_users).sum())


def test_graph_handler_split_and_datasets_cpu():
    g = synthetic.make_graph("tiny", seed=0)
    h = dh.GraphDataHandler(g.edge_index, g.num_users, g.num_items, device="cpu")
    h.set_split(g.train_idx, g.val_idx, g.test_idx)
    tr, va, te = h.get_datasets()
    assert torch.equal(tr.edge_index, g.edges("train")) and tr.num_nodes == g.num_nodes
    assert torch.equal(tr.n_id, torch.arange(g.num_nodes)) and h.get_num_users_items() == (200, 300)
    np.random.seed(1)
    h2 = dh.GraphDataHandler(g.edge_index, g.num_users, g.num_items, device="cpu")
    a, b, c = h2.get_datasets()
    e = g.edge_index.shape[1]
    assert a.edge_index.shape[1] == round(0.9 * e) and b.edge_index.shape[1] + c.edge_index.shape[1] == e - round(0.9 * e)
