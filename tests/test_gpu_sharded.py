"""GPU: the node-range sharded step on the CUDA backend.  world_size 1 always; world_size 2..N when the box has
that many GPUs (gpurun --gpus N): small shapes against the CPU oracle, BASELINE shapes (ML-25M: C3; the 10x graph:
C5, opt-in) against a float64 restatement on the device (tests/fp64_ref.py)."""
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


def test_host_shard_pipeline_same_losses_as_device_resident_steps():
    """sharded.HostShardPipeline (upload of step i+1 overlapped with step i, CSR pair rebuilt every step) against the
    same steps on a device-resident edge list with identical negatives (same torch CUDA seed).  The loss and the clip norm
    are double sums accumulated with one atomic per CTA, so two runs may differ in the last float bit -- nothing more."""
    dev = torch.device("cuda:0")
    g = synthetic.make_graph("ml100k", seed=0)
    train = g.edges("train")
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)

    def run(pipelined):
        ops = sharded.CudaOps(train.to(dev), g.num_users, g.num_items, 3)
        tr = sharded.ShardedTrainer(ops, u0.to(dev).clone(), i0.to(dev).clone())
        torch.manual_seed(11)
        if not pipelined:
            return [float(tr.step_sampled(g.num_items, use_graph=False)) for _ in range(5)], tr.gather_weights()
        pipe = sharded.HostShardPipeline(tr, train.clone().pin_memory(), None)
        return [float(pipe.step(g.num_items)) for _ in range(5)], tr.gather_weights()
    a, wa = run(False)
    b, wb = run(True)
    assert max(abs(x - y) / abs(x) for x, y in zip(a, b)) < 1e-6
    assert max_abs(wa[0], wb[0]) < 1e-7 and max_abs(wa[1], wb[1]) < 1e-7
    with pytest.raises(ValueError):
        sharded.HostShardPipeline(None, train, None)                 # pageable memory: the copy could not overlap


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


# ---------------------------------------------------------------------------------------------
# BASELINE sizes on N GPUs: forward, loss and dL/dE0 of one sharded step against float64
# ---------------------------------------------------------------------------------------------

def _full_size_worker(rank, world, port, out_dir, edges_path, nu, ni, k):
    import numpy as np
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}", device_id=dev)
    train = torch.from_numpy(np.load(edges_path)).to(torch.int64)
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    ops = sharded.CudaOps(train, nu, ni, k, device=dev)
    tr = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
    fin = tr.propagate_only().clone()
    # every rank's copy of the exchanged table is complete and bit-identical
    digest = torch.stack([fin.double().sum(), fin.double().abs().sum(), fin.view(torch.int32).long().sum().double()])
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(torch.equal(d, all_d[0]) for d in all_d)
    gen = torch.Generator().manual_seed(29)
    neg = torch.randint(0, ni, (ops.num_triplets,), generator=gen).to(dev)
    loss = float(tr.step(neg))
    own = {"segs": tr.segs, "grad": [ops.grad[rb:re].cpu() for rb, re in tr.segs], "loss": loss, "same": same,
           "p2p": ops.p2p, "multicast": ops.multicast, "shard_edges": int(ops.g.num_edges)}
    if rank == 0:
        own["final"] = fin.cpu()
    torch.save(own, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _run_full_size(tmp_path, world, shape):
    import numpy as np
    import fp64_ref
    nu, ni, _, k = synthetic.SHAPES[shape]
    g = synthetic.make_graph(shape, seed=0)
    train = g.edges("train")
    del g
    edges_path = f"/dev/shm/lgcn_test_{shape}_{os.getpid()}.npy"
    np.save(edges_path, train.numpy().astype(np.int32))
    try:
        mp.spawn(_full_size_worker, args=(world, _free_port(), str(tmp_path), edges_path, nu, ni, k), nprocs=world,
                 join=True)
    finally:
        os.remove(edges_path)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    dev = torch.device("cuda:0")
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    e0 = torch.cat([u0, i0]).to(dev).double()
    gen = torch.Generator().manual_seed(29)
    neg = torch.randint(0, ni, (int((train[0] < nu).sum()),), generator=gen).to(dev)
    loss, grad, final = fp64_ref.step_loss_and_grad(train.to(dev), e0, k, nu, neg)
    got_grad = torch.zeros(nu + ni, 64)
    for r in res:
        assert r["same"], "ranks hold different copies of the exchanged table"
        assert abs(r["loss"] - loss) < 1e-5 * abs(loss)
        for (rb, re), rows in zip(r["segs"], r["grad"]):
            got_grad[rb:re] = rows
    e_fin, e_grad = normwise(res[0]["final"], final), normwise(got_grad, grad)
    print(f"{shape} on {world} GPUs: p2p {res[0]['p2p']} multicast {res[0]['multicast']} shard edges "
          f"{[r['shard_edges'] for r in res]} of {train.shape[1]}; final normwise {e_fin:.2e}, dL/dE0 normwise {e_grad:.2e}, "
          f"loss {res[0]['loss']:.7f} vs fp64 {loss:.7f}")
    assert e_fin < TOL and e_grad < TOL


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_ml25m_step_vs_fp64(tmp_path, world):
    """BASELINE config C3 at full size on `world` GPUs."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _run_full_size(tmp_path, world, "ml25m")


@pytest.mark.skipif(os.environ.get("LGCN_RUN_C5") != "1", reason="set LGCN_RUN_C5=1 (minutes of host time)")
def test_multi_gpu_c5_10x_graph_step_vs_fp64(tmp_path):
    """BASELINE config C5 (1.6 M x 0.6 M, 225 M train edges, K = 4) on all GPUs of the box."""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    _run_full_size(tmp_path, world, "ml25m_x10")


def _score_worker(rank, world, port, out_dir):
    from lgcn_b200.utils import recommend as rec
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}", device_id=dev)
    g = synthetic.make_graph("ml1m", seed=0)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 3)
    out = rec.sharded_full_rank_eval(u0.to(dev), i0.to(dev), g.edges("train").to(dev), g.edges("test").to(dev),
                                     g.num_users, k=20)
    torch.save(out, os.path.join(out_dir, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_multi_gpu_sharded_scoring_matches_single_gpu(tmp_path, world):
    """C4 sharded over user ranges: every rank ends with the single-GPU recall / NDCG."""
    from lgcn_b200.utils import recommend as rec
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_score_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    g = synthetic.make_graph("ml1m", seed=0)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 3)
    want = rec.full_rank_eval(u0.to(dev), i0.to(dev), g.edges("train").to(dev), g.edges("test").to(dev), g.num_users, k=20)
    res = [torch.load(tmp_path / f"s{r}.pt") for r in range(world)]
    assert [r["user_range"] for r in res] == rec.user_ranges(g.num_users, world)
    for r in res:
        assert r["users"] == want["users"]
        assert abs(r["recall"] - want["recall"]) < 1e-12 and abs(r["ndcg"] - want["ndcg"]) < 1e-12
