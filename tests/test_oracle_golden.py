"""CPU: pin the oracle restatement against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py), closed-form known answers, and an independent sparse-CSR oracle."""
import numpy as np
import torch

import lgcn_b200  # noqa: F401
from lgcn_b200.data import synthetic
from oracle import pyg_restated as pyg
from oracle import reference_path as ref
from conftest import normwise, max_abs, ADAM_STEP_ATOL

TOL = 1e-5   # normwise, fp32 (BASELINE.json north_star)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_smoke_matching_golden_and_closed_form(golden):
    g = golden("smoke_matching.npz")
    u0, i0, ei = _t(g["user_w"]), _t(g["item_w"]), _t(g["edge_index"])
    uf, itf = ref.forward(u0, i0, ei, int(g["num_layers"]))
    assert torch.equal(uf, _t(g["user_final"])) and torch.equal(itf, _t(g["item_final"]))
    # closed form on the perfect matching: K=4 -> (3 U0 + 2 I0)/25, isolated -> I0/25
    assert normwise(uf, (3 * u0 + 2 * i0[:10]) / 25) < 1e-6
    assert normwise(itf[10:], i0[10:] / 25) < 1e-6
    su, si = ref.get_embeddings(u0, i0, torch.tensor([0, 1, 2]), torch.tensor([3, 4, 5, 6]))
    assert torch.equal(su, _t(g["sel_user"])) and torch.equal(si, _t(g["sel_item"]))


def _step_case(golden, name):
    g = golden(name)
    shape = str(g["shape"])
    gr = synthetic.make_graph(shape, seed=0)
    train, val = gr.edges("train"), gr.edges("val")
    assert int(train.sum()) == int(g["train_checksum"]) and int(val.sum()) == int(g["val_checksum"])
    k, s = int(g["num_layers"]), int(g["row_stride"])
    u0, i0 = synthetic.init_embeddings(gr.num_users, gr.num_items, 64, 0)
    return g, gr, train, val, k, s, u0, i0


def test_forward_loss_grads_match_reference(golden):
    for name in ["tiny_step.npz", "ml100k_step.npz"]:
        g, gr, train, val, k, s, u0, i0 = _step_case(golden, name)
        uf, itf = ref.forward(u0, i0, train, k)
        assert torch.equal(uf[::s], _t(g["user_final"])) and torch.equal(itf[::s], _t(g["item_final"]))
        user, pos = ref.triplet_users_pos(train, gr.num_users)
        assert user.numel() == int(g["num_triplets"])
        assert torch.equal(user[::s], _t(g["user"]).long()) and torch.equal(pos[::s], _t(g["pos"]).long())
        neg = _t(g["neg"]).long()
        loss, gu, gi = ref.loss_and_grads(u0, i0, train, neg, k)
        assert abs(float(loss) - float(g["loss"])) <= 1e-7 * abs(float(g["loss"])) + 1e-9
        assert normwise(gu[::s], _t(g["grad_user"])) < 1e-6 and normwise(gi[::s], _t(g["grad_item"])) < 1e-6
        # independent oracle: torch.sparse CSR matmul, and an fp64 run of the restatement
        u2, i2 = ref.forward_spmm(u0, i0, train, k)
        assert normwise(u2, uf) < TOL and normwise(i2, itf) < TOL
        u64, i64 = ref.forward(u0.double(), i0.double(), train, k)
        assert normwise(uf, u64) < TOL and normwise(itf, i64) < TOL


def test_train_epoch_and_evaluate_match_reference(golden):
    for name in ["tiny_step.npz", "ml100k_step.npz"]:
        g, gr, train, val, k, s, u0, i0 = _step_case(golden, name)
        st = ref.TrainState(u0, i0, k)
        batches = [train[:, 0::2].contiguous(), train[:, 1::2].contiguous()]
        negs = [_t(g["neg_b0"]).long(), _t(g["neg_b1"]).long()]
        ep = ref.train_epoch(st, batches, negs)
        assert abs(ep - float(g["epoch_loss"])) < 1e-6 * abs(float(g["epoch_loss"]))
        assert max_abs(st.user_w.detach()[::s], _t(g["user_w_after"])) < 2 * ADAM_STEP_ATOL
        assert max_abs(st.item_w.detach()[::s], _t(g["item_w_after"])) < 2 * ADAM_STEP_ATOL
        # evaluate(): loss on the val edges + degenerate sampled recall on LAYER-0 rows
        uw, iw = st.user_w.detach(), st.item_w.detach()
        vneg = _t(g["val_neg"]).long()
        vloss = ref.loss_from_weights(uw, iw, val, vneg, k)
        assert abs(float(vloss) - float(g["val_loss"])) < 1e-4 * abs(float(g["val_loss"]))
        user, pos = ref.triplet_users_pos(val, gr.num_users)
        draws = [_t(d).long() for d in g["recall_draws"]]
        rec = ref.compute_recall_at_k(uw[user], iw[pos], iw[vneg], draws, k=100)
        assert abs(rec - float(g["val_recall"])) < 2e-2 * float(g["val_recall"])


def test_to_undirected_split_cluster_pipeline_match_reference(golden):
    g = golden("cluster_pipeline.npz")
    # id maps + rating>=4 filter + to_undirected (dataset_handler.py:105-141)
    keep = g["csv_rating"] >= 4
    cu, cm = g["csv_user"][keep], g["csv_movie"][keep]
    _, first_u = np.unique(cu, return_index=True)
    u_order = cu[np.sort(first_u)]
    _, first_m = np.unique(cm, return_index=True)
    m_order = cm[np.sort(first_m)]
    assert np.array_equal(u_order, g["user_id_keys"]) and np.array_equal(m_order, g["movie_id_keys"])
    nu, nm = int(g["num_users"]), int(g["num_movies"])
    assert nu == len(u_order) and nm == len(m_order)
    umap = {int(k): i for i, k in enumerate(u_order)}
    mmap = {int(k): i + nu for i, k in enumerate(m_order)}
    ei = torch.tensor([[umap[int(x)] for x in cu], [mmap[int(x)] for x in cm]])
    und = pyg.to_undirected(ei)
    assert torch.equal(und, _t(g["edge_index"]).long())
    # the synthetic generator's shortcut equals to_undirected on its own pairs
    keys = (und[0, : und.shape[1] // 2] * nm + (und[1, : und.shape[1] // 2] - nu))
    assert torch.equal(synthetic.undirected_edge_index(torch.sort(keys)[0], nu, nm), und)
    # split bookkeeping (dataset_handler.py:201-233)
    val_idx, test_idx = np.sort(g["val_idx"]), np.sort(g["test_idx"])
    train_idx = np.setdiff1d(np.arange(und.shape[1]), np.concatenate([val_idx, test_idx]))
    train = und[:, _t(train_idx)]
    assert torch.equal(train, _t(g["train_edges"]).long())
    # Cluster-GCN remap given the partition vector (dataset_handler.py:273-282)
    parts = ref.cluster_batches(train, nu + nm, _t(g["cluster"]), int(g["num_parts"]))
    sizes = g["part_sizes"]
    assert [p.shape[1] for p in parts] == sizes.tolist()
    assert torch.equal(torch.cat(parts, 1), _t(g["part_edges"]).long())
    # closed form: part p == {(r,c) in train: cluster[r]==cluster[c]==p} in (r,c) order
    cl = _t(g["cluster"])
    for p, pe in enumerate(parts):
        m = (cl[train[0]] == p) & (cl[train[1]] == p)
        assert torch.equal(pe, train[:, m])


def test_recommend_from_user_matches_reference(golden):
    g = golden("cluster_pipeline.npz")
    nu, nm = int(g["num_users"]), int(g["num_movies"])
    u0, i0 = synthetic.init_embeddings(nu, nm, 64, int(g["rec_seed"]))
    uidx = int(np.where(g["user_id_keys"] == int(g["rec_user_id"]))[0][0])
    ids, vals = ref.recommend_scores(u0, i0, uidx, _t(g["rec_excluded"]), top=10)
    titles = [f"Movie {int(g['movie_id_keys'][i])}" for i in ids]
    assert titles == [str(t) for t in g["rec_titles"]]
    assert np.allclose(vals, g["rec_scores"], rtol=0, atol=1e-7)
    assert str(g["bad_error"]) == "Invalid user ID"


def test_degree_structs_consistent():
    gr = synthetic.make_graph("tiny", seed=1)
    train = gr.edges("train")
    n = gr.num_nodes
    d = ref.degree_structs(train, n)
    deg, dis, w = pyg.gcn_norm(train, n, torch.float32)
    assert torch.equal(deg.long(), d["in_deg"])
    assert int(d["ptr_in"][-1]) == train.shape[1] == int(d["ptr_out"][-1])
    # asymmetric graph: some source has in-degree 0 in a directed split
    assert torch.equal(train[0][d["eid_by_target"]], d["src_by_target"])
    assert (dis[d["in_deg"] == 0] == 0).all()


def test_full_rank_topk_and_metrics():
    gr = synthetic.make_graph("tiny", seed=2)
    u0, i0 = synthetic.init_embeddings(gr.num_users, gr.num_items, 64, 3)
    train, test = gr.edges("train"), gr.edges("test")
    um = train[:, train[0] < gr.num_users]
    excl = {u: (um[1, um[0] == u] - gr.num_users) for u in range(gr.num_users)}
    ids, vals = ref.full_rank_topk(u0, i0, excl, 20)
    for u in range(0, gr.num_users, 17):
        assert not torch.isin(ids[u], excl[u]).any()
        assert (vals[u][:-1] >= vals[u][1:]).all()
    tm = test[:, test[0] < gr.num_users]
    truth = {u: (tm[1, tm[0] == u] - gr.num_users) for u in range(gr.num_users)}
    r, n = ref.recall_ndcg_at_k(ids, truth, list(range(gr.num_users)), 20)
    assert 0.0 <= r <= 1.0 and 0.0 <= n <= 1.0
