"""Where does a C2 epoch go?  Times (a) train() over device batches (CUDA-graph replays), (b) the bare
replays back to back, (c) the end-of-epoch flush, (d) train() over host batches (staged path) split into
upload / batched build / steps, host time vs device time.  Run on the GPU box: python tools/epoch_breakdown.py"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import lgcn_b200  # noqa: E402,F401
from lgcn_b200 import _lib  # noqa: E402
from lgcn_b200.data import synthetic  # noqa: E402
from lgcn_b200.data.dataset_handler import ClusterData, ClusterLoader, Data  # noqa: E402
from lgcn_b200.models.light_gcn import LightGCN  # noqa: E402
from lgcn_b200.utils import train_test as tt  # noqa: E402

dev = torch.device("cuda:0")
g = synthetic.make_graph("ml25m", seed=0)
train = g.edges("train")
n = g.num_nodes
cluster = bench.load_partition(train, n, "ml25m")
cd = ClusterData(Data(edge_index=train.to(dev), num_nodes=n), 100, cluster=cluster)
parts = [d for d in cd.parts]
model = LightGCN(g.num_users, g.num_items, num_layers=3).to(dev)
opt = tt.FusedAdam(model)
loader = ClusterLoader(parts, shuffle=False)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(); fn(); z.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        out.append((a.elapsed_time(z), (t1 - t0) * 1e3))
    return np.median([x[0] for x in out]), np.median([x[1] for x in out])


for _ in range(3):
    tt.train(model, opt, loader, dev)
print("train() device batches, persistent kernel: device %.2f ms, host %.2f ms" % timed(lambda: tt.train(model, opt, loader, dev)))

# per-phase stamps of the persistent kernel (LGCN_EPOCH_PROF=1)
if os.environ.get("LGCN_EPOCH_PROF"):
    tt.train(model, opt, loader, dev)
    torch.cuda.synchronize()
    L = _lib.lib()
    nb = len(parts)
    ws = opt.buffers.steps_ws
    desc_bytes = L.lgcn_train_steps_workspace_bytes(nb) - 512 - 128 * nb
    prof = ws[desc_bytes + 512: desc_bytes + 512 + 128 * nb].view(torch.int64).view(nb, 16).cpu().numpy()
    dur = np.diff(prof[:, :9], axis=1) / 1e3
    names = ["fwd1 (reads e0)", "fwd2", "fwd3", "E bpr (user rows)", "bwd1+negs", "bwd2", "bwd3", "J adam+fill"]
    print("per-phase us over %d steps (median / mean / max):" % nb)
    for i, nm in enumerate(names):
        print("  %-20s %7.1f %7.1f %7.1f" % (nm, np.median(dur[:, i]), dur[:, i].mean(), dur[:, i].max()))
    tot = (prof[:, 8] - prof[:, 0]) / 1e3
    print("  step total           %7.1f %7.1f %7.1f   epoch kernel %.2f ms" % (np.median(tot), tot.mean(), tot.max(),
                                                                               (prof[-1, 8] - prof[0, 0]) / 1e6))
    hp = (prof[:-1, 13] - prof[:-1, 12]) / 1e3          # helper CTAs: prefetch + prepare of the next step
    print("  helpers (prepare)    %7.1f %7.1f %7.1f   steps where the helpers take longer than the main phases: %d"
          % (np.median(hp), hp.mean(), hp.max(), int((prof[:-1, 13] > prof[:-1, 7] + 1000 * np.median(dur[:, 7])).sum())))

if os.environ.get("QUICK"):
    sys.exit(0)
tt.EPOCH_KERNEL = False
for _ in range(3):
    tt.train(model, opt, loader, dev)
print("train() device batches, per-batch CUDA graphs: device %.2f ms, host %.2f ms" % timed(lambda: tt.train(model, opt, loader, dev)))
tt.EPOCH_KERNEL = True


def flush_only():
    opt.pending = True
    opt.flush()


tt.train(model, opt, loader, dev)
print("flush alone (nothing pending): device %.2f ms, host %.2f ms" % timed(flush_only))

# staged path
host_parts = [Data(edge_index=d.edge_index.cpu().pin_memory(), num_nodes=n) for d in parts]
for _ in range(2):
    tt.train(model, opt, host_parts, dev)
print("train() host batches (staged): device %.2f ms, host %.2f ms" % timed(lambda: tt.train(model, opt, host_parts, dev)))

eis = [d.edge_index for d in host_parts]
sizes = [int(e.shape[1]) for e in eis]
off = np.zeros(len(eis) + 1, dtype=np.int64)
np.cumsum(sizes, out=off[1:])
st = opt.stage


def upload():
    import ctypes
    ptrs = (ctypes.c_void_p * len(eis))(*[t.data_ptr() for t in eis])
    _lib.check(_lib.lib().lgcn_upload_lists(ptrs, off.ctypes.data, len(eis), st.edges.data_ptr(), _lib.stream_ptr(dev)))


print("upload 100 lists: device %.2f ms, host %.2f ms" % timed(upload))
print("batched build: device %.2f ms, host %.2f ms" % timed(
    lambda: _lib.BatchedGraphs(st.edges, off, model.num_users, model.num_items, st.arena, st.workspace)))
bg = _lib.BatchedGraphs(st.edges, off, model.num_users, model.num_items, st.arena, st.workspace)
trip = [x.num_triplets for x in bg.graphs]
neg_all = torch.randint(0, model.num_items, (sum(trip),), device=dev)
print("100 steps in one persistent launch: device %.2f ms, host %.2f ms" % timed(
    lambda: tt._launch_steps(model, opt, bg.graphs, neg_all, opt.losses.data_ptr(), 5e-3, dev)))
opt.flush()
