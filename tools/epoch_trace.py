"""Per-CTA barrier trace of the persistent step kernel (diagnostic build -DEP_TRACE=1).

Every CTA records globaltimer when it ARRIVES at (all its warps are done) and when it LEAVES each barrier of
each step.  Tells, per phase, how much is work of the median CTA, how much is waiting for the slowest CTA, and
how much is the barrier itself (last arrival -> first / last exit).

    python tools/epoch_trace.py build      (here, no GPU: builds csrc/build/variants/liblgcn_trace.so)
    python tools/epoch_trace.py            (on the GPU box)
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
VARIANT = os.path.join(REPO, "movie-recommender-system-with-gnns_b200", "csrc", "build", "variants", "liblgcn_trace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    import lgcn_b200  # noqa: F401
    from lgcn_b200 import build as _b
    print(_b.build_variant("trace", {"EP_TRACE": 1}))
    sys.exit(0)

os.environ["LGCN_LIB_PATH"] = VARIANT
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import lgcn_b200  # noqa: E402,F401
from lgcn_b200.data import synthetic  # noqa: E402
from lgcn_b200.data.dataset_handler import ClusterData, ClusterLoader, Data  # noqa: E402
from lgcn_b200.models.light_gcn import LightGCN  # noqa: E402
from lgcn_b200.utils import train_test as tt  # noqa: E402

dev = torch.device("cuda:0")
g = synthetic.make_graph("ml25m", seed=0)
train = g.edges("train")
n = g.num_nodes
K = 3
cluster = bench.load_partition(train, n, "ml25m")
cd = ClusterData(Data(edge_index=train.to(dev), num_nodes=n), 100, cluster=cluster)
parts = [d for d in cd.parts]
model = LightGCN(g.num_users, g.num_items, num_layers=K).to(dev)
opt = tt.FusedAdam(model)
loader = ClusterLoader(parts, shuffle=False)
for _ in range(3):
    tt.train(model, opt, loader, dev)
torch.cuda.synchronize()

sms = torch.cuda.get_device_properties(dev).multi_processor_count
grid = sms * 3
nb = len(parts)
trace = torch.zeros(nb, 16, grid, 2, dtype=torch.int64, device=dev)
wtrace = torch.zeros(nb, 16, grid, 8, dtype=torch.int64, device=dev)
smid = torch.zeros(2 * grid, dtype=torch.int32, device=dev)
fine = torch.zeros(nb, 16, 4096, 8, dtype=torch.int64, device=dev)
os.environ["LGCN_EPOCH_TRACE_FINE_PTR"] = str(fine.data_ptr())
os.environ["LGCN_EPOCH_TRACE_PTR"] = str(trace.data_ptr())
os.environ["LGCN_EPOCH_TRACE_WARP_PTR"] = str(wtrace.data_ptr())
os.environ["LGCN_EPOCH_TRACE_SMID_PTR"] = str(smid.data_ptr())
tt.train(model, opt, loader, dev)
torch.cuda.synchronize()
tr = trace.cpu().numpy().astype(np.float64) / 1e3          # us
# roles are taken per SM inside the kernel: put the main CTAs first, each group in the order of its role index
info = smid.cpu().numpy()
role, cidx = info[grid:] >> 16, info[grid:] & 0xffff
perm = np.concatenate([np.where(role == 0)[0][np.argsort(cidx[role == 0])], np.where(role == 1)[0][np.argsort(cidx[role == 1])]])
nmain = int((role == 0).sum())
nhelp = grid - nmain
tr = tr[:, :, perm, :]
last = 2 * K + 2                                            # slot of the end-of-step barrier
names = {1: "fwd1", 2: "fwd2", 3: "fwd3", 4: "E bpr (user rows)", 5: "bwd1+negs", 6: "bwd2", 7: "bwd3", 8: "J adam (+fill)"}
print(f"grid {grid} CTAs: {nmain} main + {nhelp} helpers; {nb} steps; us, median over steps 1..{nb - 1}")
print("%-16s %8s %8s %8s %8s | %8s %8s | %8s" % ("phase", "arr min", "arr med", "arr p90", "arr max", "rel first", "rel last", "phase"))
tot = {}
for s in range(1, last + 1):
    rows = []
    for b in range(1, nb):
        prev_exit = tr[b, s - 1, :nmain, 1].max() if s > 1 else tr[b - 1, last, :, 1].max()
        arr = tr[b, s, :nmain, 0] - prev_exit
        ex = tr[b, s, :nmain, 1]
        arr_last = tr[b, s, :nmain, 0].max()
        if s == last:                                        # everybody meets: helpers may arrive last
            arr_last = max(arr_last, tr[b, s, nmain:, 0].max())
        rows.append((arr.min(), np.median(arr), np.percentile(arr, 90), arr.max(), ex.min() - arr_last, ex.max() - arr_last,
                     ex.max() - prev_exit))
    r = np.median(np.array(rows), axis=0)
    tot[s] = r[6]
    print("%-16s %8.2f %8.2f %8.2f %8.2f | %8.2f %8.2f | %8.2f" % (names.get(s, str(s)), *r))
print("sum of phases %.1f us" % sum(tot.values()))
# helpers: when do they arrive at the end-of-step barrier, relative to the step start and to the main CTAs' arrival
h_rows = []
for b in range(1, nb - 1):
    start = tr[b - 1, last, :, 1].max()
    h_arr = tr[b, last, nmain:, 0]
    m_arr = tr[b, last, :nmain, 0]
    fill = tr[b, last + 1, :nmain, 0] - m_arr
    h_rows.append((np.median(h_arr) - start, h_arr.max() - start, m_arr.max() - start, np.median(fill), fill.max()))
h = np.array(h_rows)
print("helpers arrive at the end-of-step barrier: median CTA %.1f us, last CTA %.1f us after the step start "
      "(main CTAs' last arrival %.1f us); helpers are last in %d of %d steps" %
      (np.median(h[:, 0]), np.median(h[:, 1]), np.median(h[:, 2]), int((h[:, 1] > h[:, 2]).sum()), len(h)))
hh = []
for b in range(1, nb - 1):
    start = tr[b - 1, last, :, 1].max()
    a1, a2, a3 = tr[b, 10, nmain:, 0] - start, tr[b, 11, nmain:, 0] - start, tr[b, last, nmain:, 0] - start
    hh.append((np.median(a1), a1.max(), np.median(a2), a2.max(), np.median(a3), a3.max()))
hh = np.median(np.array(hh), axis=0)
print("helper CTAs, us after the step start (median CTA / last CTA): replay of the next step's active rows done %.1f / %.1f, "
      "negative list built %.1f / %.1f, list rows replayed %.1f / %.1f" % tuple(hh))
print("cache fill after arriving: median CTA %.2f us, slowest CTA %.2f us" % (np.median(h[:, 3]), np.median(h[:, 4])))
# per-CTA arrival order: is it always the same CTAs that are late?  (SM / die placement)
late = np.zeros(nmain)
for s in range(1, last):
    for b in range(1, nb):
        a_ = tr[b, s, :nmain, 0]
        late += (a_ - a_.min())
late /= (last - 1) * (nb - 1)
order = np.argsort(late)
print("mean lateness per main CTA (us): min %.2f  median %.2f  max %.2f; latest CTAs %s" %
      (late.min(), np.median(late), late.max(), order[-8:].tolist()))

# ---- per-warp arrivals: who is late, and what did it have to do? -------------------------------------------------
wt = wtrace.cpu().numpy().astype(np.float64)[:, :, perm, :] / 1e3
sm = info[:grid][perm]
helpers_on_sm = np.bincount(sm[nmain:], minlength=sms)
print("SMs by number of helper CTAs:", np.bincount(helpers_on_sm).tolist(), " main CTAs per SM:", np.bincount(np.bincount(sm[:nmain], minlength=sms)).tolist())
nw = nmain * 8
for slot, lst in ((2, "in_tasks"), (6, "out_tasks")):
    stats = []          # (lateness, edges of the warp's tasks, number of tasks, helper CTAs on its SM, warp index)
    for b in range(1, nb):
        gr = model.graph(parts[b].edge_index)
        tk = getattr(gr, lst).view(-1, 8).cpu().numpy()
        nt = gr.c.n_in_tasks if lst == "in_tasks" else gr.c.n_out_tasks
        if nt > 2 * nw:
            continue                                        # the hub cluster: its own regime
        ln = (tk[:nt, 2] - tk[:nt, 1])
        prev_exit = tr[b, slot - 1, :nmain, 1].max()
        for c in range(nmain):
            for wv in range(8):
                gwi = wv * nmain + c
                mine = ln[gwi:nt:nw]
                stats.append((wt[b, slot, c, wv] - prev_exit, mine.sum(), len(mine), mine.max() if len(mine) else 0,
                              helpers_on_sm[sm[c]], wv))
    st = np.array(stats)
    print(f"slot {slot} ({names[slot]}): warp arrival after the phase start, us: median %.2f  p90 %.2f  p99 %.2f  max(median over steps) ~%.2f"
          % (np.median(st[:, 0]), np.percentile(st[:, 0], 90), np.percentile(st[:, 0], 99), np.percentile(st[:, 0], 99.95)))
    for label, col, bins in (("tasks of the warp", 2, [0, 1, 2, 3]), ("longest task (edges)", 3, [0, 1, 5, 17, 33, 65]),
                             ("helper CTAs on the SM", 4, [1, 2, 3]), ("warp index in CTA", 5, list(range(8)))):
        out = []
        for i, lo in enumerate(bins):
            hi = bins[i + 1] if i + 1 < len(bins) else 1 << 30
            sel = (st[:, col] >= lo) & (st[:, col] < hi)
            if sel.any():
                out.append("[%d..): n=%d med %.2f p99 %.2f" % (lo, sel.sum(), np.median(st[sel, 0]), np.percentile(st[sel, 0], 99)))
        print("   by %-22s %s" % (label, " | ".join(out)))

# ---- clock64 stamps inside a forward phase (slot 2 = fwd2): cycles from the warp leaving the previous barrier ----
fn = fine.cpu().numpy().astype(np.float64)
rows = []
for b in range(1, nb):
    gr = model.graph(parts[b].edge_index)
    nt = gr.c.n_in_tasks
    if nt > 2 * nw:
        continue
    tk = gr.in_tasks.view(-1, 8).cpu().numpy()
    ln = tk[:nt, 2] - tk[:nt, 1]
    for gwi in range(min(nw, 4096)):
        f = fn[b, 2, gwi]
        if f[7] == 0 or f[0] == 0:
            continue
        ntask = len(ln[gwi:nt:nw])
        first = ln[gwi] if gwi < nt else 0
        base = f[7]                                       # left the barrier that ended fwd1
        rows.append((ntask, first, f[0] - base, f[1] - base, f[2] - base, f[3] - base, f[4] - base, f[5] - base, f[6] - base))
r = np.array(rows)
print("fwd2, SM cycles after the warp left the previous barrier (median): enter run_tasks / descriptor read / gather done / "
      "shuffles done / epilogue done / loop done / at barrier")
for label, sel in (("no task", r[:, 0] == 0), ("1 task, <=4 edges", (r[:, 0] == 1) & (r[:, 1] <= 4)),
                   ("1 task, 5-16 edges", (r[:, 0] == 1) & (r[:, 1] > 4) & (r[:, 1] <= 16)),
                   ("1 task, 17-32", (r[:, 0] == 1) & (r[:, 1] > 16) & (r[:, 1] <= 32)),
                   ("1 task, 33-64", (r[:, 0] == 1) & (r[:, 1] > 32)), ("2 tasks", r[:, 0] == 2)):
    if sel.any():
        print("   %-20s n=%6d  " % (label, sel.sum()) + "  ".join("%6.0f" % np.median(r[sel, i]) for i in range(2, 9)))
