"""GPU: K0b (lgcn_graph_build_batched) gives, for every list, exactly the arrays lgcn_graph_build gives
for that list alone, and an epoch over HOST batches (staged upload + batched build) applies exactly the
optimiser steps the per-batch device path applies."""
import numpy as np
import pytest
import torch

import lgcn_b200  # noqa: F401
from lgcn_b200 import _lib
from lgcn_b200.data import synthetic
from lgcn_b200.data.dataset_handler import Data
from lgcn_b200.models.light_gcn import LightGCN
from lgcn_b200.utils import train_test as tt
from conftest import ADAM_STEP_ATOL, max_abs, normwise

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None

ARRAYS = ["in_ptr", "out_ptr", "in_nbr", "in_trip", "out_nbr", "out_trip", "dis", "active", "in_tasks", "out_tasks",
          "active_list"]


def _hash_batches(g, parts):
    train = g.edges("train")
    cl = synthetic.hash_partition(g.num_nodes, parts)
    keep = cl[train[0]] == cl[train[1]]
    e, p = train[:, keep], cl[train[0]][keep]
    return [e[:, p == q].contiguous() for q in range(parts)]


def _pack(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([x.shape[1] for x in lists], out=off[1:])
    flat = torch.cat([x.contiguous().reshape(-1) for x in lists]) if lists else torch.zeros(0, dtype=torch.int64)
    return flat.to(DEV), off


def _same_as_single(view, ei, nu, ni):
    ref = _lib.Graph(ei.to(DEV), nu, ni)
    for f in ("num_edges", "num_triplets", "n_in_tasks", "n_out_tasks", "n_in_user_tasks", "n_out_user_tasks",
              "n_in_slots", "n_out_slots", "num_active", "row_split", "num_nodes", "num_users"):
        assert getattr(view.c, f) == getattr(ref.c, f), f
    for name in ARRAYS:
        a = view.array(name)
        b = getattr(ref, name).reshape(-1)[: a.numel()]
        assert torch.equal(a, b), name


def test_batched_build_bit_identical_to_single_builds():
    g = synthetic.make_graph("ml100k", seed=0)
    lists = _hash_batches(g, 7)
    lists.insert(2, torch.zeros(2, 0, dtype=torch.int64))                     # an empty list in the middle
    lists.append(torch.tensor([[3, g.num_users + 5], [g.num_users + 5, 3]]))   # two edges
    lists.append(g.edges("train"))                                            # the whole train graph (split rows)
    lists.append(g.edges("val")[:, g.edges("val")[0] >= g.num_users])         # movie->user only: P = 0
    edges, off = _pack(lists)
    bg = _lib.BatchedGraphs(edges, off, g.num_users, g.num_items)
    assert len(bg) == len(lists)
    for view, ei in zip(bg.graphs, lists):
        _same_as_single(view, ei, g.num_users, g.num_items)
    assert bg.graphs[-1].num_triplets == 0 and bg.graphs[2].num_edges == 0
    # arena / workspace reuse gives the same result
    bg2 = _lib.BatchedGraphs(edges, off, g.num_users, g.num_items, bg.arena, bg.workspace)
    assert bg2.arena.data_ptr() == bg.arena.data_ptr()
    _same_as_single(bg2.graphs[0], lists[0], g.num_users, g.num_items)


def test_batched_build_large_list_uses_large_row_split():
    g = synthetic.make_graph_custom(3000, 2500, 700_000, seed=2)              # 1.26 M train edges >= LGCN_SMALL_GRAPH
    train = g.edges("train")
    assert train.shape[1] >= (1 << 20)
    lists = [train, g.edges("val")]
    edges, off = _pack(lists)
    bg = _lib.BatchedGraphs(edges, off, g.num_users, g.num_items)
    assert bg.graphs[0].c.row_split == 512 and bg.graphs[1].c.row_split == 64
    for view, ei in zip(bg.graphs, lists):
        _same_as_single(view, ei, g.num_users, g.num_items)


def test_batched_build_rejects_bad_input():
    edges = torch.tensor([0, 1, 5, 6], dtype=torch.int64, device=DEV)          # 1 -> 6 ok, 0 -> 5 ok ...
    _lib.BatchedGraphs(edges, [0, 2], 4, 4)
    bad = torch.tensor([0, 1, 2, 6], dtype=torch.int64, device=DEV)            # 0 -> 2 joins two users
    with pytest.raises(_lib.LgcnError):
        _lib.BatchedGraphs(bad, [0, 2], 4, 4)
    with pytest.raises(_lib.LgcnError):
        _lib.BatchedGraphs(edges, [0, 3], 4, 4)                                # offsets beyond the buffer
    with pytest.raises(_lib.LgcnError):
        _lib.BatchedGraphs(edges, [1, 2], 4, 4)                                # edge_off[0] != 0


def _model(g, k=3):
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    m = LightGCN(g.num_users, g.num_items, num_layers=k).to(DEV)
    with torch.no_grad():
        m.user_embedding.weight.copy_(u0)
        m.item_embedding.weight.copy_(i0)
    return m


@pytest.mark.parametrize("pinned", [True, False], ids=["pinned", "pageable"])
def test_staged_host_epoch_equals_per_batch_steps(pinned):
    g = synthetic.make_graph("ml100k", seed=1)
    lists = [x for x in _hash_batches(g, 24)]                                 # small enough for sparse steps
    lists.insert(4, torch.zeros(2, 0, dtype=torch.int64))
    lists.append(g.edges("train"))                                            # a dense (non-sparse) step
    host = [Data(edge_index=x.pin_memory() if pinned else x, num_nodes=g.num_nodes) for x in lists]

    m1, m2 = _model(g), _model(g)
    o1, o2 = tt.FusedAdam(m1), tt.FusedAdam(m2)
    epochs = 2
    torch.cuda.manual_seed(1234)
    losses1 = [tt.train(m1, o1, host, DEV) for _ in range(epochs)]

    # the same negatives the staged path drew: one randint per run of consecutive sparse-eligible batches,
    # one per dense batch (utils/train_test.py::_run_staged)
    torch.cuda.manual_seed(1234)
    live = [x for x in lists if x.shape[1] > 0]
    dev_lists = [x.to(DEV) for x in live]
    graphs = [m2.graph(x) for x in dev_lists]
    groups, cur = [], []
    for i, gr in enumerate(graphs):
        if tt.sparse_step_pays(gr):
            cur.append(i)
        else:
            if cur:
                groups.append(cur)
            groups.append([i])
            cur = []
    if cur:
        groups.append(cur)
    assert any(len(gp) > 1 for gp in groups) and any(not tt.sparse_step_pays(graphs[gp[0]]) for gp in groups)
    losses2 = []
    for _ in range(epochs):
        acc, wsum = 0.0, 0
        for gp in groups:
            trip = [graphs[i].num_triplets for i in gp]
            neg_all = torch.randint(0, g.num_items, (sum(trip),), device=DEV)
            o = 0
            for i, p in zip(gp, trip):
                loss = tt.train_step(m2, o2, dev_lists[i], neg_all[o:o + p], sparse=tt.sparse_step_pays(graphs[i]))
                o += p
                acc += float(loss.item()) * live[i].shape[1]
                wsum += live[i].shape[1]
        o2.flush()
        losses2.append(acc / wsum)
    o1.flush()
    # same kernels on bit-identical graphs; the only run-to-run freedom is the order of the float atomics
    # that scatter the negative-item gradient (bpr.cu pass A), amplified by Adam (tests/conftest.py)
    steps = epochs * len(live)
    assert max_abs(m1.user_embedding.weight, m2.user_embedding.weight) <= ADAM_STEP_ATOL * steps
    assert max_abs(m1.item_embedding.weight, m2.item_embedding.weight) <= ADAM_STEP_ATOL * steps
    assert normwise(o1.exp_avg, o2.exp_avg) < 1e-4
    assert np.allclose(losses1, losses2, rtol=2e-5, atol=0), (losses1, losses2)
    assert int(o1.step_count) == epochs * len(live)
