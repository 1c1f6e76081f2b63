"""Where does a Cluster-GCN batch step spend its time?  Event-timed loops over single C-ABI calls on
the median and the largest ML-25M-shaped batch (development aid; results summarised in DESIGN.md)."""
import os
import sys
from ctypes import byref

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from bench import NUM_PARTS, load_partition  # noqa: E402
from lgcn_b200 import _lib  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402
from lgcn_b200.data.dataset_handler import ClusterData, Data  # noqa: E402
from lgcn_b200.models.light_gcn import LightGCN  # noqa: E402
from lgcn_b200.utils import train_test as tt  # noqa: E402

dev = torch.device("cuda:0")
g = synthetic.make_graph("ml25m", seed=0)
train = g.edges("train")
cluster = load_partition(train, g.num_nodes, "ml25m")
cd = ClusterData(Data(edge_index=train.to(dev), num_nodes=g.num_nodes), NUM_PARTS, cluster=cluster)
sizes = np.array([d.edge_index.shape[1] for d in cd.parts])
order = np.argsort(sizes)
model = LightGCN(g.num_users, g.num_items, num_layers=3).to(dev)
opt = tt.FusedAdam(model)
L = _lib.lib()
s = _lib.stream_ptr(dev)
uw, iw = model.user_embedding.weight, model.item_embedding.weight
b = opt.buffers


def timed(fn, iters=200, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    z.record()
    torch.cuda.synchronize()
    return a.elapsed_time(z) / iters * 1e3      # us


for name, idx in (("median", order[len(order) // 2]), ("p90", order[int(0.9 * len(order))]), ("largest", order[-1])):
    ei = cd.parts[int(idx)].edge_index
    G = model.graph(ei)
    neg = torch.randint(0, g.num_items, (G.num_triplets,), device=dev)
    b.ensure_triplets(G.num_triplets)
    reg = 2.0 * 5e-3 / (64.0 * G.num_triplets)
    loss = torch.empty(1, device=dev)
    k = 3
    res = {"E": int(ei.shape[1]), "P": G.num_triplets, "active": G.num_active, "in_tasks": G.c.n_in_tasks,
           "slots": G.c.n_in_slots}
    res["launch_floor(step_begin)"] = timed(lambda: L.lgcn_step_begin(byref(opt.c), b.accum.data_ptr(), s))
    res["fwd_dense(4 launches)"] = timed(lambda: L.lgcn_propagate_fwd(G.ref, uw.data_ptr(), iw.data_ptr(), k, b.final_emb.data_ptr(),
                                                                    b.rnorm.data_ptr(), b.work.data_ptr(), b.work.numel() * 4, s))
    res["bpr_dense"] = timed(lambda: L.lgcn_bpr_fwd_bwd(G.ref, b.final_emb.data_ptr(), b.rnorm.data_ptr(), neg.data_ptr(),
                                                       b.grad_final.data_ptr(), b.neg_count.data_ptr(), b.trip_scratch.data_ptr(),
                                                       b.accum.data_ptr(), s))
    res["bwd_dense(4 launches)"] = timed(lambda: L.lgcn_propagate_bwd(G.ref, b.grad_final.data_ptr(), k, uw.data_ptr(), iw.data_ptr(),
                                                                    b.neg_count.data_ptr(), reg, b.grad_e0.data_ptr(), b.accum.data_ptr(),
                                                                    b.work.data_ptr(), b.work.numel() * 4, s))
    res["clip_adam_dense"] = timed(lambda: L.lgcn_clip_adam(byref(opt.c), uw.data_ptr(), iw.data_ptr(), model.num_users, model.num_items,
                                                           b.grad_e0.data_ptr(), b.accum.data_ptr(), G.num_triplets, 5e-3, loss.data_ptr(), s))
    res["spmm_single_layer"] = timed(lambda: L.lgcn_spmm(G.ref, b.grad_final.data_ptr(), b.grad_e0.data_ptr(), 0, s))
    res["dense_step_eager"] = timed(lambda: tt.train_step(model, opt, ei, neg, loss), iters=50)
    opt.flush()
    res["sparse_step_eager"] = timed(lambda: tt.train_step(model, opt, ei, neg, loss, sparse=True), iters=100)
    opt.flush()
    opt.dirty = False
    b.grad_final.zero_(); b.neg_count.zero_()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        tt._launch_step(model, opt, G, neg, loss, 5e-3, True)
    opt.pending = True
    res["sparse_step_graph"] = timed(cg.replay, iters=200)
    res["flush"] = timed(lambda: (setattr(opt, "pending", True), opt.flush()), iters=20)
    opt.flush()
    b.grad_final.zero_(); b.neg_count.zero_()
    cg2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg2):
        tt._launch_step(model, opt, G, neg, loss, 5e-3, False)
    res["dense_step_graph"] = timed(cg2.replay, iters=50)
    print(name, {k_: (round(v, 1) if isinstance(v, float) else v) for k_, v in res.items()})
