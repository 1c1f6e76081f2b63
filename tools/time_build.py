"""Where does the C3 e2e step go?  Upload of the 22.5 M-edge list, lgcn_graph_build (K0) and its host side."""
import os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa
from lgcn_b200 import _lib
from lgcn_b200.data import synthetic
dev = torch.device("cuda:0")
g = synthetic.make_graph("ml25m", seed=0)
train = g.edges("train")
host = train.pin_memory()

def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    for _ in range(reps):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record(); r = fn(); z.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print("%-28s device %.2f ms  wall %.2f ms" % (name, a.elapsed_time(z), (t1 - t0) * 1e3))
    return r

ei = timed("upload 360 MB", lambda: host.to(dev, non_blocking=True))
timed("Graph() total", lambda: _lib.Graph(ei, g.num_users, g.num_items))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    gr = _lib.Graph(ei, g.num_users, g.num_items); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
