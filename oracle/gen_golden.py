"""ORACLE tooling (test infrastructure) -- generate tests/golden/*.npz by running the
UNMODIFIED reference modules from /root/reference.

Runs only in the build container (needs /root/reference; the GPU box has no such path).
The reference's third-party imports that are not installable offline are satisfied by stub
modules backed by ``oracle/pyg_restated.py``:

    torch_geometric.nn.LGConv, torch_geometric.data.Data,
    torch_geometric.loader.{ClusterData, DataLoader}, torch_geometric.utils.to_undirected,
    memory_profiler.profile (unused decorator import, data/dataset_handler.py:14)

Everything else executed here is the reference's own code:
    models/light_gcn.py      LightGCN.__init__/forward/get_embeddings
    utils/helpers.py         get_triplets_indices / sample_negative
    utils/train_test.py      bpr_loss, compute_embeddings, train, evaluate, compute_recall_at_k
    data/dataset_handler.py  MovieLensDataHandler (CSV -> id maps -> to_undirected -> split ->
                             ClusterData remap -> DataLoader)
    utils/recommend.py       recommend_from_user (visualizations stubbed: plotly/umap absent)

Usage:  python -m oracle.gen_golden          (from the repo root)
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

from oracle import pyg_restated as pyg  # noqa: E402
import lgcn_b200  # noqa: E402,F401  (import alias for the hyphenated package dir)
from lgcn_b200.data import synthetic  # noqa: E402  (input generator only; no kernels)


def _install_stubs() -> None:
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_nn.LGConv = pyg.LGConv
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data = pyg.Data
    tg_loader = types.ModuleType("torch_geometric.loader")
    tg_loader.ClusterData = pyg.ClusterData
    tg_loader.DataLoader = pyg.DataLoader
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_utils.to_undirected = pyg.to_undirected
    tg.nn, tg.data, tg.loader, tg.utils = tg_nn, tg_data, tg_loader, tg_utils
    mp = types.ModuleType("memory_profiler")
    mp.profile = lambda f: f
    vis = types.ModuleType("visualizations")
    vis.plot_recommendations = lambda *a, **k: None
    vis.analyze_user_recommendations = lambda *a, **k: None
    vis.plot_histories = lambda *a, **k: None
    for name, mod in [("torch_geometric", tg), ("torch_geometric.nn", tg_nn),
                      ("torch_geometric.data", tg_data), ("torch_geometric.loader", tg_loader),
                      ("torch_geometric.utils", tg_utils), ("memory_profiler", mp),
                      ("visualizations", vis)]:
        sys.modules[name] = mod


def _import_reference():
    _install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "utils"))  # train_test.py does `from helpers import`
    from models.light_gcn import LightGCN
    import helpers
    import train_test
    import recommend
    from data.dataset_handler import MovieLensDataHandler
    return LightGCN, helpers, train_test, recommend, MovieLensDataHandler


class _Batch:
    """What the reference's loop needs from a batch: .edge_index and .to(device)
    (utils/train_test.py:87,98,120)."""

    def __init__(self, edge_index):
        self.edge_index = edge_index

    def to(self, device):
        return self


def _set_weights(model, seed):
    u, i = synthetic.init_embeddings(model.num_users, model.num_items, model.dim_h, seed)
    with torch.no_grad():
        model.user_embedding.weight.copy_(u)
        model.item_embedding.weight.copy_(i)
    return u, i


def gen_smoke_matching(LightGCN):
    """models/light_gcn.py:66-89 smoke graph (perfect matching u_i <-> item_i, K=4)."""
    ei = torch.tensor([list(range(20)), list(range(10, 20)) + list(range(10))], dtype=torch.long)
    model = LightGCN(10, 15)
    u0, i0 = _set_weights(model, 7)
    with torch.no_grad():
        uf, itf = model.forward(ei)
        su, si = model.get_embeddings(user_indices=torch.tensor([0, 1, 2]),
                                      item_indices=torch.tensor([3, 4, 5, 6]))
    np.savez(os.path.join(OUT, "smoke_matching.npz"), edge_index=ei.numpy(), user_w=u0.numpy(),
             item_w=i0.numpy(), user_final=uf.numpy(), item_final=itf.numpy(),
             sel_user=su.numpy(), sel_item=si.numpy(), num_layers=4)


def gen_step(LightGCN, helpers, train_test, shape, name, row_stride):
    """Forward, triplets, loss, autograd grads, then the reference's own train() for one
    epoch over two batches, then evaluate() on the val edges."""
    g = synthetic.make_graph(shape, seed=0)
    k = synthetic.SHAPES[shape][3]
    train = g.edges("train")
    val = g.edges("val")
    dev = torch.device("cpu")
    model = LightGCN(g.num_users, g.num_items, num_layers=k, dim_h=64)
    u0, i0 = _set_weights(model, 0)

    # forward + loss + grads, negatives from a recorded seed
    torch.manual_seed(1234)
    embs = train_test.compute_embeddings(model, _Batch(train), dev)
    torch.manual_seed(1234)
    user, pos, neg = helpers.get_triplets_indices(train, g.num_users, g.num_items, dev)
    loss = train_test.bpr_loss(*embs)
    model.zero_grad()
    loss.backward()
    gu, gi = model.user_embedding.weight.grad.clone(), model.item_embedding.weight.grad.clone()
    with torch.no_grad():
        uf, itf = model(train)

    # one epoch of the reference's train() over two batches (even / odd edge positions, so
    # both contain user->movie edges; a batch without any yields NaN)
    batches = [train[:, 0::2].contiguous(), train[:, 1::2].contiguous()]
    model2 = LightGCN(g.num_users, g.num_items, num_layers=k, dim_h=64)
    _set_weights(model2, 0)
    opt = torch.optim.Adam(model2.parameters(), lr=0.001)       # utils/train_test.py:236
    torch.manual_seed(4321)
    epoch_loss = train_test.train(model2, opt, [_Batch(b) for b in batches], dev)
    torch.manual_seed(4321)                                      # replay the negative stream
    negs = [helpers.get_triplets_indices(b, g.num_users, g.num_items, dev)[2] for b in batches]
    # evaluate(): val loss + the degenerate sampled recall; record the np.random draws
    np.random.seed(99)
    torch.manual_seed(777)
    val_loss, recall = train_test.evaluate(model2, _Batch(val), dev, top_k=100)
    torch.manual_seed(777)
    val_neg = helpers.get_triplets_indices(val, g.num_users, g.num_items, dev)[2]
    np.random.seed(99)
    p_val = int((val[0] < g.num_users).sum())
    draws = np.stack([np.random.choice(p_val, 100, replace=False) for _ in range(10)])

    s = row_stride
    np.savez_compressed(os.path.join(OUT, name),
             shape=shape, num_layers=k, num_users=g.num_users, num_items=g.num_items,
             train_checksum=int(train.sum()), val_checksum=int(val.sum()),
             user=user.numpy()[::s].astype(np.int32), pos=pos.numpy()[::s].astype(np.int32), neg=neg.numpy().astype(np.int32), num_triplets=user.numel(),
             row_stride=s,
             user_final=uf.numpy()[::s], item_final=itf.numpy()[::s], loss=float(loss.detach()),
             grad_user=gu.numpy()[::s], grad_item=gi.numpy()[::s],
             grad_norm=float(torch.sqrt(gu.pow(2).sum() + gi.pow(2).sum())),
             neg_b0=negs[0].numpy().astype(np.int32), neg_b1=negs[1].numpy().astype(np.int32), epoch_loss=float(epoch_loss),
             user_w_after=model2.user_embedding.weight.detach().numpy()[::s],
             item_w_after=model2.item_embedding.weight.detach().numpy()[::s],
             val_neg=val_neg.numpy().astype(np.int32), val_loss=float(val_loss), val_recall=float(recall),
             recall_draws=draws)
    print(name, "loss", float(loss), "epoch_loss", float(epoch_loss), "val", float(val_loss), recall)


def gen_cluster_pipeline(MovieLensDataHandler, recommend, LightGCN):
    """data/dataset_handler.py end to end on a synthetic ratings.csv, plus
    utils/recommend.py::recommend_from_user on the resulting handler."""
    import pandas as pd
    u_n, i_n, cnt = 300, 400, 12_000
    keys = synthetic.make_interactions(u_n, i_n, cnt, seed=3)
    gen = torch.Generator().manual_seed(5)
    order = torch.randperm(cnt, generator=gen)                   # CSV row order (id-map order)
    u = (keys // i_n)[order].numpy()
    m = (keys % i_n)[order].numpy()
    user_ids = u * 7 + 11                                        # non-contiguous raw ids
    movie_ids = m * 3 + 5
    # add low ratings that the >=4 filter (dataset_handler.py:106) must drop
    low_u = np.arange(50) * 7 + 11
    low_m = np.arange(50) * 3 + 5 + 3 * i_n
    ratings = pd.DataFrame({
        "userId": np.concatenate([user_ids, low_u]),
        "movieId": np.concatenate([movie_ids, low_m]),
        "rating": np.concatenate([np.where(np.arange(cnt) % 2 == 0, 4.0, 5.0), np.full(50, 3.5)]),
        "timestamp": 0,
    })
    all_movies = np.unique(np.concatenate([movie_ids, low_m]))
    movies = pd.DataFrame({"movieId": all_movies, "title": [f"Movie {x}" for x in all_movies],
                           "genres": "x"})
    num_parts = 8
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            os.makedirs("data/movielens-25m")
            rp, mp_ = "data/movielens-25m/ratings.csv", "data/movielens-25m/movies.csv"
            ratings.to_csv(rp, index=False)
            movies.to_csv(mp_, index=False)
            np.random.seed(2024)                                 # split is otherwise unseeded (:167-168)
            torch.manual_seed(0)
            h = MovieLensDataHandler(rp, mp_)
            loader, val_ds, test_ds = h.get_data_training(num_train_clusters=num_parts)
            train_ds, _, _ = h.get_datasets()                    # reloads the persisted split
            val_idx = np.load("data/indexes/val_indices.npy")
            test_idx = np.load("data/indexes/test_indices.npy")
            parts = [b.edge_index for b in loader.dataset]       # unshuffled order p = 0..P-1
            n = h.num_users + h.num_movies
            # recover the METIS vector the stub computed, to make it an explicit input
            se, _ = pyg.sort_edge_index(train_ds.edge_index, n)
            cluster = pyg.metis_partition(pyg.index2ptr(se[0], n), se[1], num_parts)
            # recommend_from_user on a seeded model
            model = LightGCN(h.num_users, h.num_movies)
            _set_weights(model, 21)
            uid = int(user_ids[0])
            uidx = h.user_id_map[uid]
            excl = train_ds.edge_index[1, train_ds.edge_index[0, :] == uidx] - h.num_users   # recommend.py:141-142
            rec = recommend.recommend_from_user(model, uid, h, excl)
            bad = recommend.recommend_from_user(model, -1, h, None)
        finally:
            os.chdir(cwd)
    sizes = np.array([p.shape[1] for p in parts])
    cat = torch.cat(parts, dim=1).numpy() if len(parts) else np.zeros((2, 0), np.int64)
    np.savez_compressed(os.path.join(OUT, "cluster_pipeline.npz"),
             csv_user=ratings["userId"].values, csv_movie=ratings["movieId"].values,
             csv_rating=ratings["rating"].values,
             num_users=h.num_users, num_movies=h.num_movies, edge_index=h.edge_index.numpy().astype(np.int32),
             val_idx=val_idx, test_idx=test_idx, train_edges=train_ds.edge_index.numpy().astype(np.int32),
             cluster=cluster.numpy(), num_parts=num_parts, part_sizes=sizes, part_edges=cat.astype(np.int32),
             user_id_keys=np.array(list(h.user_id_map.keys())), user_id_vals=np.array(list(h.user_id_map.values())),
             movie_id_keys=np.array(list(h.movie_id_map.keys())), movie_id_vals=np.array(list(h.movie_id_map.values())),
             rec_user_id=uid, rec_titles=np.array([r["title"] for r in rec["recommendations"]]),
             rec_scores=np.array([r["score"] for r in rec["recommendations"]]),
             rec_excluded=excl.numpy(), rec_seed=21, bad_error=bad["error"])
    print("cluster_pipeline parts", sizes.tolist(), "rec", [r["title"] for r in rec["recommendations"]][:3])


def gen_bench_partition():
    """METIS vector for the ML-25M-shaped train graph (BASELINE config C2) so bench.py does
    not spend ~75 s in METIS on every run.  Stored with a checksum of the edges it was
    computed on; bench.py recomputes if the checksum does not match."""
    g = synthetic.make_graph("ml25m", seed=0)
    train = g.edges("train")
    n = g.num_nodes
    se, _ = pyg.sort_edge_index(train, n)
    cluster = pyg.metis_partition(pyg.index2ptr(se[0], n), se[1], 100)
    np.savez_compressed(os.path.join(OUT, "ml25m_seed0_metis100.npz"), cluster=cluster.numpy().astype(np.int8),
                        train_checksum=int((train[0] * 31 + train[1]).sum()), num_parts=100)
    print("ml25m partition stored")


def main():
    os.makedirs(OUT, exist_ok=True)
    LightGCN, helpers, train_test, recommend, Handler = _import_reference()
    gen_smoke_matching(LightGCN)
    gen_step(LightGCN, helpers, train_test, "tiny", "tiny_step.npz", 1)
    gen_step(LightGCN, helpers, train_test, "ml100k", "ml100k_step.npz", 8)
    gen_cluster_pipeline(Handler, recommend, LightGCN)
    if "--with-bench-partition" in sys.argv:
        gen_bench_partition()


if __name__ == "__main__":
    main()
