"""Where does the sharded full-graph step spend its time?  Event-timed stages on every rank, max over ranks.

    python tools/time_sharded.py                                  (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/time_sharded.py

LGCN_BENCH_SHAPE=ml25m|ml25m_x10 picks the graph.  Output: one line of microseconds per stage (rank 0).
"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from lgcn_b200 import _lib, sharded  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
shape = os.environ.get("LGCN_BENCH_SHAPE", "ml25m")
k = synthetic.SHAPES[shape][3]
cache = os.environ.get("LGCN_EDGE_CACHE")             # A/B runs of library variants: generate the graph once
if cache and world == 1 and os.path.exists(cache):
    import numpy as np
    nu, ni = synthetic.SHAPES[shape][0], synthetic.SHAPES[shape][1]
    train = torch.from_numpy(np.load(cache)).to(torch.int64)
else:
    nu, ni, train, _ = synthetic.shared_train_edges(shape, local, dist.barrier if world > 1 else (lambda: None))
    if cache and world == 1:
        import numpy as np
        np.save(cache, train.numpy().astype(np.int32))


class g:                                   # noqa: N801  (the few graph facts used below)
    num_users, num_items = nu, ni


ops = sharded.CudaOps(train, nu, ni, k, device=dev)
u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
t = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
torch.manual_seed(0)
neg = torch.randint(0, g.num_items, (ops.num_triplets,), device=dev)
first_loss = float(t.step(neg))


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t.comm.barrier()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    z.record()
    torch.cuda.synchronize()
    v = torch.tensor([a.elapsed_time(z) / iters * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v)


def synced(fn):
    """a stage followed by the exchange barrier it needs (stages run on all ranks in lock step)"""
    def run():
        fn()
        if ops.p2p:
            ops.peer_barrier()
    return run


res = {}
if ops.p2p:
    res["peer_barrier"] = timed(ops.peer_barrier)
    res["allreduce_accum"] = timed(ops.allreduce_accum)
res["prescale"] = timed(synced(ops.prescale))
for layer in range(1, k + 1):
    res[f"fwd_layer{layer}"] = timed(synced(lambda layer=layer: ops.fwd_layer(layer)))
res["bpr_owner"] = timed(synced(lambda: ops.bpr(neg)))
for j in range(1, k + 1):
    res[f"bwd_layer{j}"] = timed(synced(lambda j=j: ops.bwd_layer(j, 5e-3)))
res["clip_adam"] = timed(lambda: ops.clip_adam(5e-3))
res["randint"] = timed(lambda: torch.randint(0, g.num_items, (ops.num_triplets,), device=dev))
res["propagate_only"] = timed(t.propagate_only)
res["step_eager"] = timed(lambda: t.step(neg), iters=10)
for _ in range(6):
    t.step_sampled(g.num_items, use_graph=True)
res["step_graph"] = timed(lambda: t.step_sampled(g.num_items, use_graph=True), iters=10)
if rank == 0:
    print(f"world {world} shape {shape} p2p {ops.p2p} multicast {ops.multicast} graph "
          f"{'yes' if getattr(t, '_graph', None) is not None else 'no: ' + str(getattr(t, '_graph_error', ''))}",
          {key: round(v, 1) for key, v in res.items()})
    print("lib", os.path.basename(_lib.LIB_PATH), "first step loss", repr(first_loss))
    print("shard edges", ops.g.num_edges, "of", ops.E, "local tasks in/out", ops.local.n_in_tasks, ops.local.n_out_tasks)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
