"""GPU: the node-range sharded step on the CUDA backend.  world_size 1 always; world_size 2..N over
NCCL when the box has that many GPUs (gpurun --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lgcn_b200  # noqa: F401
from conftest import ADAM_STEP_ATOL, max_abs, normwise
from lgcn_b200 import sharded
from lgcn_b200.data import synthetic
from oracle import reference_path as ref

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _negs(p, ni, steps):
    gen = torch.Generator().manual_seed(23)
    return [torch.randint(0, ni, (p,), generator=gen) for _ in range(steps)]


def _oracle(shape, k, steps):
    g = synthetic.make_graph(shape, seed=0)
    train = g.edges("train")
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    st = ref.TrainState(u0, i0, k)
    p = int((train[0] < g.num_users).sum())
    losses = [st.step(train, n) for n in _negs(p, g.num_items, steps)]
    return g, train, u0, i0, st, losses


@pytest.mark.parametrize("shape,k", [("tiny", 3), ("ml100k", 3), ("ml1m", 2)])
def test_world_one_sharded_step_matches_oracle_and_fused_step(shape, k):
    from lgcn_b200.models.light_gcn import LightGCN
    from lgcn_b200.utils import train_test as tt
    dev = torch.device("cuda:0")
    steps = 2
    g, train, u0, i0, st, want = _oracle(shape, k, steps)
    ops = sharded.CudaOps(train.to(dev), g.num_users, g.num_items, k)
    tr = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev))
    negs = _negs(ops.num_triplets, g.num_items, steps)
    got = [float(tr.step(n.to(dev))) for n in negs]
    assert max(abs(a - b) / abs(b) for a, b in zip(got, want)) < 1e-4
    uw, iw = tr.gather_weights()
    assert max_abs(uw, st.user_w.detach()) < steps * ADAM_STEP_ATOL and max_abs(iw, st.item_w.detach()) < steps * ADAM_STEP_ATOL
    # same numbers as the fused single-call step (different layer-1 kernel: pre-scaled table)
    m = LightGCN(g.num_users, g.num_items, num_layers=k).to(dev)
    with torch.no_grad():
        m.user_embedding.weight.copy_(u0)
        m.item_embedding.weight.copy_(i0)
    opt = tt.FusedAdam(m)
    fused = [float(tt.train_step(m, opt, train.to(dev), n.to(dev))) for n in negs]
    assert max(abs(a - b) / abs(b) for a, b in zip(got, fused)) < 1e-5
    fin = tr.propagate_only()
    uf, itf = ref.forward(uw.cpu().double(), iw.cpu().double(), train, k)
    assert normwise(fin, torch.cat([uf, itf])) < TOL


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, shape, k, steps, p2p):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}", device_id=dev)
    g = synthetic.make_graph(shape, seed=0)
    train = g.edges("train")
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    ops = sharded.CudaOps(train.to(dev), g.num_users, g.num_items, k, p2p=p2p)
    tr = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
    losses = [float(tr.step(n.to(dev))) for n in _negs(ops.num_triplets, g.num_items, steps)]
    uw, iw = tr.gather_weights()
    fin = tr.propagate_only()
    torch.save({"losses": losses, "uw": uw.cpu(), "iw": iw.cpu(), "final": fin.cpu(), "p2p": ops.p2p,
                "multicast": getattr(ops, "multicast", False), "p2p_error": ops.p2p_error},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [True, False], ids=["fused-p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_sharded_step_matches_oracle(tmp_path, world, p2p):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    shape, k, steps = "ml100k", 3, 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), shape, k, steps, p2p), nprocs=world, join=True)
    g, train, u0, i0, st, want = _oracle(shape, k, steps)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    print("p2p", res[0]["p2p"], "multicast", res[0]["multicast"], "error", res[0]["p2p_error"])
    for r in res:
        assert max(abs(a - b) / abs(b) for a, b in zip(r["losses"], want)) < 1e-4
        assert max_abs(r["uw"], st.user_w.detach()) < steps * ADAM_STEP_ATOL
        assert max_abs(r["iw"], st.item_w.detach()) < steps * ADAM_STEP_ATOL
        assert torch.equal(r["uw"], res[0]["uw"]) and torch.equal(r["final"], res[0]["final"])
    uf, itf = ref.forward(res[0]["uw"].double(), res[0]["iw"].double(), train, k)
    assert normwise(res[0]["final"], torch.cat([uf, itf])) < TOL
