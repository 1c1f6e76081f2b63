"""CPU, world_size 2 over gloo: the node-range sharded training step (lgcn_b200/sharded.py) equals
the unsharded oracle.  The compute backend is the torch test double in tests/sharded_cpu_ops.py; the
orchestration, ownership plan and collectives under test are the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lgcn_b200  # noqa: F401
from lgcn_b200 import sharded
from lgcn_b200.data import synthetic
from oracle import reference_path as ref
from sharded_cpu_ops import TorchOps


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _negs(p, ni, steps):
    gen = torch.Generator().manual_seed(17)
    return [torch.randint(0, ni, (p,), generator=gen) for _ in range(steps)]


def _worker(rank, world, port, out_dir, k, steps):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    torch.set_num_threads(2)
    g = synthetic.make_graph("tiny", seed=0)
    train = g.edges("train")
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    ops = TorchOps(train, g.num_users, g.num_items, k)
    tr = sharded.ShardedTrainer(ops, u0.double().clone(), i0.double().clone(), sharded.Comm())
    losses = [float(tr.step(n)) for n in _negs(ops.num_triplets, g.num_items, steps)]
    fin = tr.propagate_only().clone()
    uw, iw = tr.gather_weights()
    torch.save({"losses": losses, "uw": uw, "iw": iw, "final": fin, "segs": tr.segs,
                "own_stale": float((tr.user_w - uw).abs().max())}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [3, 2])
def test_two_rank_sharded_step_equals_unsharded_oracle(tmp_path, k):
    steps = 2
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), k, steps), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in (0, 1))
    g = synthetic.make_graph("tiny", seed=0)
    train = g.edges("train")
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    st = ref.TrainState(u0.double(), i0.double(), k)
    p = int((train[0] < g.num_users).sum())
    want = [st.step(train, n) for n in _negs(p, g.num_items, steps)]
    for r in (r0, r1):
        assert max(abs(a - b) for a, b in zip(r["losses"], want)) < 1e-9
        assert float((r["uw"] - st.user_w.detach()).abs().max()) < 1e-7      # Adam amplifies fp64 noise too
        assert float((r["iw"] - st.item_w.detach()).abs().max()) < 1e-7
    assert torch.equal(r0["uw"], r1["uw"]) and torch.equal(r0["final"], r1["final"])
    uf, itf = ref.forward(st.user_w.detach(), st.item_w.detach(), train, k)
    assert float((r0["final"] - torch.cat([uf, itf])).abs().max()) < 1e-9
    # ownership: the two ranks tile the user range and the item range exactly, both non-trivially
    (u0b, u0e), (i0b, i0e) = r0["segs"]
    (u1b, u1e), (i1b, i1e) = r1["segs"]
    assert (u0b, u1e, i0b, i1e) == (0, g.num_users, g.num_users, g.num_nodes) and u0e == u1b and i0e == i1b
    assert 0 < u0e < g.num_users and g.num_users < i0e < g.num_nodes
    assert r0["own_stale"] > 0            # each rank only stepped its own rows; gather_weights repaired the rest


def test_world_size_one_plan_and_balance():
    g = synthetic.make_graph("ml100k", seed=0)
    train = g.edges("train")
    ind, outd = torch.bincount(train[1], minlength=g.num_nodes), torch.bincount(train[0], minlength=g.num_nodes)
    for world in (1, 2, 4, 8):
        plan = sharded.ShardPlan.build(ind, outd, g.num_users, world)
        assert plan.user_ptr[0] == 0 and plan.user_ptr[-1] == g.num_users
        assert plan.item_ptr[0] == g.num_users and plan.item_ptr[-1] == g.num_nodes
        w = (ind + outd + 8).double()
        loads = [float(w[a:b].sum() + w[c:d].sum()) for (a, b), (c, d) in (plan.segments(r) for r in range(world))]
        assert max(loads) / (sum(loads) / world) < 1.15              # edge-balanced within 15 %
    assert sharded.balanced_boundaries(torch.zeros(0), 3) == [0, 0, 0, 0]
