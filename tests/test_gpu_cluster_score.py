"""GPU parity: Cluster-GCN extraction (K4, bit-exact) and fused scoring + mask + top-k (K5)."""
import os

import numpy as np
import pytest
import torch

import lgcn_b200  # noqa: F401
from conftest import normwise
from lgcn_b200 import _lib
from lgcn_b200.data import dataset_handler as dh
from lgcn_b200.data import synthetic
from lgcn_b200.models.light_gcn import LightGCN
from lgcn_b200.utils import recommend as rec
from oracle import pyg_restated as pyg
from oracle import reference_path as ref

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


def _t(a):
    return torch.from_numpy(np.asarray(a))


# ---------------------------------------------------------------------------------------------
# K4
# ---------------------------------------------------------------------------------------------

def test_cluster_extract_matches_reference_pipeline_golden(golden):
    g = golden("cluster_pipeline.npz")
    n = int(g["num_users"]) + int(g["num_movies"])
    train = _t(g["train_edges"]).long()
    edges, part_ptr = dh.cluster_extract(train.to(DEV), n, _t(g["cluster"]), int(g["num_parts"]))
    assert torch.equal(part_ptr.cpu()[1:] - part_ptr.cpu()[:-1], _t(g["part_sizes"]).long())
    assert torch.equal(edges.cpu(), _t(g["part_edges"]).long())
    # to_undirected on the device == the reference's edge_index
    raw = _t(g["edge_index"]).long()
    half = raw[:, : raw.shape[1] // 2]
    assert torch.equal(dh.to_undirected(half.to(DEV)).cpu(), raw)


@pytest.mark.parametrize("shape,parts,shuffle", [("tiny", 7, False), ("ml100k", 100, False), ("ml100k", 16, True),
                                                 ("ml1m", 100, False)])
def test_cluster_extract_bit_exact_vs_oracle(shape, parts, shuffle):
    g = synthetic.make_graph(shape, seed=1)
    train = g.edges("train")
    if shuffle:
        train = train[:, torch.randperm(train.shape[1], generator=torch.Generator().manual_seed(0))].contiguous()
    cluster = synthetic.hash_partition(g.num_nodes, parts)
    cluster[cluster == 3] = 2                                  # an EMPTY part (no nodes)
    want = ref.cluster_batches(train, g.num_nodes, cluster, parts)
    cd = dh.ClusterData(dh.Data(edge_index=train.to(DEV), num_nodes=g.num_nodes), parts, cluster=cluster)
    assert len(cd) == parts
    for p in range(parts):
        assert torch.equal(cd[p].edge_index.cpu(), want[p]), p
        assert cd[p].num_nodes == g.num_nodes
    assert cd[3].edge_index.shape[1] == 0
    # properties at any size: kept edges are exactly the intra-cluster ones, each once
    kept = torch.cat([d.edge_index for d in cd], 1).cpu()
    m = cluster[train[0]] == cluster[train[1]]
    key = lambda e: torch.sort(e[0] * g.num_nodes + e[1])[0]
    assert torch.equal(key(kept), key(train[:, m]))


def test_metis_partition_same_call_as_oracle_and_loader_contract():
    g = synthetic.make_graph("ml100k", seed=2)
    train = g.edges("train")
    n = g.num_nodes
    part = dh.metis_partition(train, n, 10)
    se, _ = pyg.sort_edge_index(train, n)
    want = pyg.metis_partition(pyg.index2ptr(se[0], n), se[1], 10)
    assert torch.equal(part, want)
    h = dh.GraphDataHandler(g.edge_index, g.num_users, g.num_items, DEV)
    h.set_split(g.train_idx, g.val_idx, g.test_idx)
    loader, val, test = h.get_data_training(num_train_clusters=10)
    assert len(loader) == 10 and val.edge_index.shape[1] == g.val_idx.numel()
    a = [b.edge_index.data_ptr() for b in loader]
    b = [b.edge_index.data_ptr() for b in loader]
    assert sorted(a) == sorted(b)                               # same tensors every epoch (cache hits)
    want = ref.cluster_batches(train, n, part, 10)
    got = {d.edge_index.shape[1] for d in loader.dataset}
    assert got == {w.shape[1] for w in want}


def test_movielens_handler_csv_pipeline_matches_reference_fixture(golden, tmp_path, monkeypatch):
    """The product's MovieLensDataHandler end to end on the CSV the reference itself ingested
    (oracle/gen_golden.py::gen_cluster_pipeline; data/dataset_handler.py:98-141,160-199,256-288): rating filter, id maps
    in order of first appearance, to_undirected (device kernel), the seeded 90/5/5 split, METIS, the cluster batches."""
    import pandas as pd
    g = golden("cluster_pipeline.npz")
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/movielens-25m")
    rp, mp = "data/movielens-25m/ratings.csv", "data/movielens-25m/movies.csv"
    pd.DataFrame({"userId": g["csv_user"], "movieId": g["csv_movie"], "rating": g["csv_rating"], "timestamp": 0}
                 ).to_csv(rp, index=False)
    all_movies = np.unique(g["csv_movie"])
    pd.DataFrame({"movieId": all_movies, "title": [f"Movie {x}" for x in all_movies], "genres": "x"}).to_csv(mp, index=False)
    np.random.seed(2024)
    h = dh.MovieLensDataHandler(rp, mp, device=DEV)
    assert (h.num_users, h.num_movies) == (int(g["num_users"]), int(g["num_movies"]))
    assert h.get_num_users_items() == (int(g["num_users"]), int(g["num_movies"]))
    assert list(h.user_id_map.keys()) == g["user_id_keys"].tolist() and list(h.user_id_map.values()) == g["user_id_vals"].tolist()
    assert list(h.movie_id_map.keys()) == g["movie_id_keys"].tolist() and list(h.movie_id_map.values()) == g["movie_id_vals"].tolist()
    assert h.id_movie_map[h.movie_id_map[int(g["movie_id_keys"][5])]] == int(g["movie_id_keys"][5])
    assert torch.equal(h.edge_index.cpu(), _t(g["edge_index"]).long())
    parts = int(g["num_parts"])
    loader, val, test = h.get_data_training(num_train_clusters=parts)          # METIS runs here
    assert np.array_equal(np.load("data/indexes/val_indices.npy"), g["val_idx"])
    assert np.array_equal(np.load("data/indexes/test_indices.npy"), g["test_idx"])
    train, _, _ = h.get_datasets()                                              # reloads the persisted split
    assert torch.equal(train.edge_index.cpu(), _t(g["train_edges"]).long())
    assert torch.equal(val.edge_index.cpu(), h.edge_index.cpu()[:, _t(g["val_idx"])])
    assert torch.equal(test.edge_index.cpu(), h.edge_index.cpu()[:, _t(g["test_idx"])])
    sizes = [b.edge_index.shape[1] for b in loader.dataset]
    assert sizes == g["part_sizes"].tolist()
    assert torch.equal(torch.cat([b.edge_index for b in loader.dataset], 1).cpu(), _t(g["part_edges"]).long())
    assert all(b.num_nodes == h.num_users + h.num_movies for b in loader.dataset)


def test_to_undirected_kernel_edge_cases():
    # duplicates in the input, both directions already present, a self loop, non-bipartite ids, explicit num_nodes
    ei = torch.tensor([[0, 0, 3, 5, 2, 4, 4, 1], [3, 3, 0, 5, 7, 1, 1, 4]], dtype=torch.long)
    got = dh.to_undirected(ei.to(DEV)).cpu()
    assert torch.equal(got, pyg.to_undirected(ei))
    assert torch.equal(dh.to_undirected(ei.to(DEV), num_nodes=12).cpu(), pyg.to_undirected(ei, 12))
    assert dh.to_undirected(torch.empty(2, 0, dtype=torch.long, device=DEV)).shape == (2, 0)
    with pytest.raises(_lib.LgcnError):
        dh.to_undirected(ei.to(DEV), num_nodes=6)                  # id 7 outside [0, 6)
    # a random multigraph with many duplicates
    gen = torch.Generator().manual_seed(3)
    big = torch.randint(0, 3000, (2, 400_000), generator=gen)
    assert torch.equal(dh.to_undirected(big.to(DEV)).cpu(), pyg.to_undirected(big))
    # the ML-25M-shaped list: the kernel reproduces the generator's closed form (25 M directed edges)
    g = synthetic.make_graph("ml25m", seed=0)
    half = g.edge_index[:, : g.edge_index.shape[1] // 2]
    assert torch.equal(dh.to_undirected(half.to(DEV)).cpu(), g.edge_index)


@pytest.mark.parametrize("shape,parts", [("ml100k", 16), ("ml1m", 100)])
def test_gpu_partitioner_vote_kernel_and_partition_match_torch_restatement(shape, parts):
    """f4: lgcn_label_vote against tests/partition_ref.py row by row, and the whole partition (device orchestration +
    kernel) against the same orchestration on the CPU with the torch vote -- bit-exact; then ClusterData with it."""
    import math
    from partition_ref import torch_vote
    from lgcn_b200.data import partition_gpu as pg
    g = synthetic.make_graph(shape, seed=0)
    tr, n = g.edges("train"), g.num_nodes
    ptr, nbr = pg.csr_by_source(tr.to(DEV), n)
    lab0 = pg.hash_labels(n, parts, DEV)
    for b, e in ((0, n), (g.num_users, n), (0, g.num_users)):
        got = pg.cuda_vote(ptr, nbr, lab0, b, e, parts)
        want = torch_vote(ptr.cpu(), nbr.cpu(), lab0.cpu(), b, e, parts)
        for a, w in zip(got, want):
            assert torch.equal(a.cpu(), w)
    lab, st = pg.partition(tr.to(DEV), n, parts, g.num_users)
    ref_lab, ref_st = pg.partition(tr, n, parts, g.num_users, vote=torch_vote)
    assert torch.equal(lab.cpu(), ref_lab) and st == ref_st
    assert st["max_part"] <= math.ceil(n / parts * 1.03)
    cd = dh.ClusterData(dh.Data(edge_index=tr.to(DEV), num_nodes=n), parts, partitioner="gpu", num_users=g.num_users)
    assert torch.equal(cd.cluster.cpu(), ref_lab)
    kept = sum(d.edge_index.shape[1] for d in cd)
    assert kept == st["intra_edges"]
    want_batches = ref.cluster_batches(tr, n, ref_lab, parts)
    assert all(torch.equal(cd[p].edge_index.cpu(), want_batches[p]) for p in range(parts))


# ---------------------------------------------------------------------------------------------
# K5
# ---------------------------------------------------------------------------------------------

def _check_topk(ids, vals, oids, ovals, tol):
    """ids must equal the oracle's wherever the oracle's neighbouring score gaps exceed tol; the
    score vectors must agree within tol everywhere."""
    assert float((vals.double() - ovals.double()).abs().max()) <= tol
    gap_prev = torch.ones_like(ovals, dtype=torch.bool)
    gap_prev[:, 1:] = (ovals[:, :-1] - ovals[:, 1:]) > 2 * tol
    gap_next = torch.ones_like(gap_prev)
    gap_next[:, :-1] = gap_prev[:, 1:]
    # the last slot also needs a gap to the first item that did NOT make the list: checked by value
    safe = gap_prev & gap_next
    safe[:, -1] = False
    assert torch.equal(ids[safe].long(), oids[safe])
    return float(safe.float().mean())


ALGOS = [(rec.SCORE_FFMA, False), (rec.SCORE_TENSOR, True), (rec.SCORE_TENSOR, False)]
ALGO_IDS = ["ffma", "tcgen05-tma-packed", "tcgen05-producer-warps"]


@pytest.mark.parametrize("algo,pack", ALGOS, ids=ALGO_IDS)
@pytest.mark.parametrize("shape,k,normalize", [("tiny", 20, True), ("ml100k", 20, True), ("ml100k", 10, False),
                                               ("ml100k", 100, True), ("ml100k", 50, True), ("ml1m", 20, True)])
def test_score_topk_vs_bruteforce_oracle(shape, k, normalize, algo, pack):
    if algo == rec.SCORE_TENSOR and k > 32:
        pytest.skip("tensor-core kernel keeps k <= 32")
    g = synthetic.make_graph(shape, seed=0)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 5)
    train = g.edges("train")
    ptr, idx = rec.exclusion_csr(train.to(DEV), g.num_users)
    ids, vals = rec.score_topk(u0.to(DEV), i0.to(DEV), k, normalize, ptr, idx, algo=algo, pack_items=pack)
    um = train[:, train[0] < g.num_users]
    order = torch.sort(um[0], stable=True)[1]
    cnt = torch.bincount(um[0], minlength=g.num_users)
    off = torch.zeros(g.num_users + 1, dtype=torch.long)
    off[1:] = torch.cumsum(cnt, 0)
    items = (um[1][order] - g.num_users)
    excl = {u: items[off[u]:off[u + 1]] for u in range(g.num_users)}
    assert torch.equal(ptr.cpu(), off)
    oids, ovals = ref.full_rank_topk(u0.double(), i0.double(), excl, k, normalize=normalize)
    scale = 1.0 if normalize else float(ovals.abs().max())
    frac = _check_topk(ids.cpu(), vals.cpu(), oids, ovals, 1e-5 * scale)
    assert frac > 0.8          # share of list positions whose neighbouring score gaps exceed 2*tol
    for u in range(0, g.num_users, 37):                         # never a masked train item
        assert not torch.isin(ids[u].cpu().long(), excl[u]).any()
    assert (vals[:, :-1] >= vals[:, 1:]).all()


@pytest.mark.parametrize("algo,pack", ALGOS, ids=ALGO_IDS)
def test_score_topk_user_range_no_exclusion_and_short_lists(algo, pack):
    gen = torch.Generator().manual_seed(3)
    u = torch.randn(300, 64, generator=gen)
    it = torch.randn(37, 64, generator=gen)                     # fewer items than one tile
    ids, vals = rec.score_topk(u.to(DEV), it.to(DEV), 20, False, u_begin=130, u_end=263, algo=algo, pack_items=pack)
    s = u[130:263].double() @ it.double().t()
    ov, oi = torch.topk(s, 20, dim=1)
    assert ids.shape == (133, 20)
    _check_topk(ids.cpu(), vals.cpu(), oi, ov, 1e-5 * float(ov.abs().max()))
    # k larger than the number of admissible items -> padded with (-1, -inf)
    ptr = torch.zeros(301, dtype=torch.int64)
    ptr[1:] = 30
    ex = torch.arange(30, dtype=torch.int32)
    ids, vals = rec.score_topk(u.to(DEV), it.to(DEV), 20, False, ptr.to(DEV), ex.to(DEV), 0, 1, algo=algo, pack_items=pack)
    assert (ids[0, :7] >= 30).all() and (ids[0, 7:] == -1).all() and torch.isinf(vals[0, 7:]).all()


def test_recommend_from_user_matches_reference_golden(golden):
    import pandas as pd
    g = golden("cluster_pipeline.npz")
    nu, nm = int(g["num_users"]), int(g["num_movies"])

    class Handler:
        user_id_map = {int(k): int(v) for k, v in zip(g["user_id_keys"], g["user_id_vals"])}
        movie_id_map = {int(k): int(v) for k, v in zip(g["movie_id_keys"], g["movie_id_vals"])}
        movies = pd.DataFrame({"movieId": g["movie_id_keys"], "title": [f"Movie {int(x)}" for x in g["movie_id_keys"]]})

    u0, i0 = synthetic.init_embeddings(nu, nm, 64, int(g["rec_seed"]))
    m = LightGCN(nu, nm).to(DEV)
    with torch.no_grad():
        m.user_embedding.weight.copy_(u0)
        m.item_embedding.weight.copy_(i0)
    out = rec.recommend_from_user(m, int(g["rec_user_id"]), Handler, _t(g["rec_excluded"]))
    assert [r["title"] for r in out["recommendations"]] == [str(t) for t in g["rec_titles"]]
    assert np.allclose([r["score"] for r in out["recommendations"]], g["rec_scores"], rtol=0, atol=1e-5)
    assert rec.recommend_from_user(m, -1, Handler, None) == {"error": str(g["bad_error"])}
    assert rec.recommend_from_movie(m, -1, Handler, None) == {"error": "Invalid movie ID"}
    top = rec.recommend_from_movie(m, int(g["movie_id_keys"][0]), Handler, None)["top_users"]
    s = torch.nn.functional.normalize(u0.double()) @ torch.nn.functional.normalize(i0.double())[0]
    assert Handler.user_id_map[top[0]["user_id"]] == int(torch.argmax(s))


def test_full_rank_eval_metrics_vs_oracle():
    g = synthetic.make_graph("ml100k", seed=0)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 8)
    train, test = g.edges("train"), g.edges("test")
    out = rec.full_rank_eval(u0.to(DEV), i0.to(DEV), train.to(DEV), test.to(DEV), g.num_users, k=20)
    um = train[:, train[0] < g.num_users]
    excl = {u: (um[1, um[0] == u] - g.num_users) for u in range(g.num_users)}
    ids, _ = ref.full_rank_topk(u0.double(), i0.double(), excl, 20)
    tm = test[:, test[0] < g.num_users]
    truth = {u: (tm[1, tm[0] == u] - g.num_users) for u in range(g.num_users)}
    r, nd = ref.recall_ndcg_at_k(ids, truth, list(range(g.num_users)), 20)
    assert abs(out["recall"] - r) < 2e-3 and abs(out["ndcg"] - nd) < 2e-3


def test_full_rank_eval_user_ranges_sum_to_the_whole():
    """SURVEY sec. 8e sharded C4: per-range partial sums (what each rank computes) add up to the unsharded metrics."""
    g = synthetic.make_graph("ml1m", seed=0)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 3)
    ue, ie = u0.to(DEV), i0.to(DEV)
    tr, te = g.edges("train").to(DEV), g.edges("test").to(DEV)
    whole = rec.full_rank_eval(ue, ie, tr, te, g.num_users, k=20)
    acc = torch.zeros(3, dtype=torch.float64, device=DEV)
    for lo, hi in rec.user_ranges(g.num_users, 3):
        part = rec.full_rank_eval(ue, ie, tr, te, g.num_users, k=20, u_begin=lo, u_end=hi)
        acc += torch.tensor([part["recall"] * part["users"], part["ndcg"] * part["users"], part["users"]],
                            dtype=torch.float64, device=DEV)
    assert int(acc[2]) == whole["users"]
    assert abs(float(acc[0] / acc[2]) - whole["recall"]) < 1e-12 and abs(float(acc[1] / acc[2]) - whole["ndcg"]) < 1e-12
    one = rec.sharded_full_rank_eval(ue, ie, tr, te, g.num_users, k=20)          # no process group: world 1
    assert one["recall"] == whole["recall"] and one["user_range"] == (0, g.num_users)
