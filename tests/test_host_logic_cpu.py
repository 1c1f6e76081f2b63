"""CPU: host-side logic that needs no device -- generator contract, loader semantics, graph cache
keying, undirected/ exclusion helpers, state_dict layout."""
import os

import numpy as np
import pytest
import torch

import lgcn_b200  # noqa: F401
from lgcn_b200._lib import LgcnError
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
    with pytest.raises(LgcnError):                                    # the product's to_undirected is a CUDA kernel
        dh.to_undirected(ei[:, :half])
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
    assert a.edge_index.shape[1] == int(0.9 * e) and b.edge_index.shape[1] + c.edge_index.shape[1] == e - int(0.9 * e)


@pytest.mark.parametrize("n", [20, 31, 1237, 24000])
def test_shuffle_split_consumes_the_stream_like_sklearn(n):
    """dataset_handler.py:167-168: the product's restated split == sklearn's train_test_split, index for index, on
    the same numpy global seed (both calls of the reference: train_size=0.9, then test_size=0.5 of the rest)."""
    from sklearn.model_selection import train_test_split
    np.random.seed(77 + n)
    want_tr, want_vt = train_test_split(np.arange(n), train_size=0.9, shuffle=True)
    want_va, want_te = train_test_split(want_vt, test_size=0.5, shuffle=True)
    np.random.seed(77 + n)
    tr, vt = dh.shuffle_split(n, train_size=0.9)
    a, b = dh.shuffle_split(len(vt), test_size=0.5)
    assert np.array_equal(tr, want_tr) and np.array_equal(vt, want_vt)
    assert np.array_equal(vt[a], want_va) and np.array_equal(vt[b], want_te)


def test_handler_split_matches_reference_fixture(tmp_path, monkeypatch):
    """The reference's own split (np.random.seed(2024), oracle/gen_golden.py::gen_cluster_pipeline) against the
    product's handler on the same edge list: val / test indices, train edges, and the .npy persist + reload path
    (dataset_handler.py:160-176, :203-253)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cluster_pipeline.npz"))
    ei = torch.from_numpy(z["edge_index"].astype(np.int64))
    nu, nm = int(z["num_users"]), int(z["num_movies"])
    h = dh.GraphDataHandler(ei, nu, nm, device="cpu")
    np.random.seed(2024)
    tr, va, te = h.get_datasets()
    assert np.array_equal(h._split[1], z["val_idx"]) and np.array_equal(h._split[2], z["test_idx"])
    assert torch.equal(tr.edge_index, torch.from_numpy(z["train_edges"].astype(np.int64)))
    assert torch.equal(va.edge_index, ei[:, torch.from_numpy(z["val_idx"])])
    # persist + reload through MovieLensDataHandler._indices (cwd-relative "data/indexes" like the reference)
    monkeypatch.chdir(tmp_path)
    m = dh.MovieLensDataHandler.__new__(dh.MovieLensDataHandler)
    dh.GraphDataHandler.__init__(m, ei, nu, nm, device="cpu")
    np.random.seed(2024)
    first = m._indices(0.9)
    assert os.path.exists("data/indexes/val_indices.npy") and os.path.exists("data/indexes/test_indices.npy")
    assert np.array_equal(np.load("data/indexes/val_indices.npy"), z["val_idx"])
    m2 = dh.MovieLensDataHandler.__new__(dh.MovieLensDataHandler)
    dh.GraphDataHandler.__init__(m2, ei, nu, nm, device="cpu")
    np.random.seed(5)                                                  # a different stream: the files decide
    again = m2._indices(0.9)
    for x, y in zip(first, again):
        assert np.array_equal(x, y)
    assert np.array_equal(again[0], np.setdiff1d(np.arange(ei.shape[1]), np.concatenate([z["val_idx"], z["test_idx"]])))


def test_user_ranges_for_sharded_scoring():
    for nu, world in [(162_541, 8), (162_541, 1), (1000, 4), (100, 8), (128, 2)]:
        r = rec.user_ranges(nu, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == nu
        assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
        assert all(lo % 128 == 0 for lo, _ in r)                     # CTA tiles are not split across ranks
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(s for s in sizes) <= 256 or nu < 128 * world


def test_partitioner_orchestration_balanced_deterministic_and_better_than_hash():
    """data/partition_gpu.py with the torch restatement of the voting kernel plugged in (tests/partition_ref.py): every
    node labelled, parts within the capacity, same result twice, more intra-part edges than the hash start."""
    import math
    from partition_ref import torch_vote
    from lgcn_b200.data import partition_gpu as pg
    g = synthetic.make_graph("ml100k", seed=0)
    tr, n, parts = g.edges("train"), g.num_nodes, 16
    lab, st = pg.partition(tr, n, parts, g.num_users, vote=torch_vote)
    lab2, _ = pg.partition(tr, n, parts, g.num_users, vote=torch_vote)
    assert torch.equal(lab, lab2) and lab.shape == (n,) and int(lab.min()) >= 0 and int(lab.max()) < parts
    assert st["max_part"] <= math.ceil(n / parts * 1.03)
    h = pg.hash_labels(n, parts, "cpu")
    assert st["intra_edges"] > 1.5 * int((h[tr[0]] == h[tr[1]]).sum())
    assert st["intra_edges"] == max(st["starts"].values())
    # one-sided graphs (no bipartite split) run too
    lab3, st3 = pg.partition(tr, n, parts, None, vote=torch_vote)
    assert st3["max_part"] <= math.ceil(n / parts * 1.03) and int(lab3.max()) < parts
