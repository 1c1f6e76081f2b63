"""Development aid (torchrun): where does the sharded step spend its time?"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ["NCCL_DEBUG"] = "WARN"
import lgcn_b200  # noqa: E402,F401
from lgcn_b200 import sharded  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
g = synthetic.make_graph("ml25m", seed=0)
tr = g.edges("train").to(dev)
ops = sharded.CudaOps(tr, g.num_users, g.num_items, 3)
u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
t = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
neg = torch.randint(0, g.num_items, (ops.num_triplets,), device=dev)


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    z.record()
    torch.cuda.synchronize()
    return a.elapsed_time(z) / iters * 1e3


res = {}
if ops.p2p:
    res["peer_barrier"] = timed(ops.peer_barrier)
res["allreduce_G_57MB"] = timed(lambda: t.comm.allreduce(ops.G))
res["allreduce_negcount"] = timed(lambda: t.comm.allreduce(ops.neg_count))
res["allreduce_accum"] = timed(lambda: t.comm.allreduce(ops.accum))
res["prescale_both"] = timed(lambda: [ops.prescale(rb, re) for rb, re in t.segs])
res["fwd_layer2_kernel"] = timed(lambda: [ops.fwd_layer(2, rb, re) for rb, re in t.layer_segs])
res["fwd_layer3_kernel"] = timed(lambda: [ops.fwd_layer(3, rb, re) for rb, re in t.layer_segs])
res["bwd_layer1_kernel"] = timed(lambda: [ops.bwd_layer(1, rb, re, 5e-3) for rb, re in t.layer_segs])
res["bpr_range"] = timed(lambda: ops.bpr(neg, *t.segs[0]))
res["clip_adam_both"] = timed(lambda: [ops.clip_adam(rb, re, 5e-3) for rb, re in t.segs])
res["propagate_only"] = timed(t.propagate_only)
res["step"] = timed(lambda: t.step(neg), iters=10)
if rank == 0:
    print(world, "p2p", ops.p2p, "mc", getattr(ops, "multicast", None), {k: round(v, 1) for k, v in res.items()})
dist.barrier()
dist.destroy_process_group()
