#!/usr/bin/env python
"""bench.py -- LightGCN hot-path benchmark (contract in the task statement / DESIGN.md sec. Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3]

N = 1  workload C2 (BASELINE.json configs[1]): Cluster-GCN training on the synthetic ML-25M-shaped
       graph (162,541 users x 59,047 movies, 25.0 M directed edges, 90 % train, 100 METIS parts,
       K = 3, dim 64).  One STEP = one training epoch = 100 cluster-batch iterations of
       utils/train_test.py:86-101 (forward, BPR loss, backward, clip, Adam).
N > 1  workload C3: full-graph training step on the same graph, node-range sharded (see
       lgcn_b200/sharded.py); launched under torchrun, one rank per GPU.
metric  lightgcn_train_edges_per_s = directed batch-graph edges consumed per second by full training
        steps (whole job); ms_per_step is the epoch (N=1) / full-graph step (N>1) time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import lgcn_b200  # noqa: E402,F401
from lgcn_b200.data import synthetic  # noqa: E402

METRIC, UNIT = "lightgcn_train_edges_per_s", "edges/s"
NUM_PARTS = 100


# --------------------------------------------------------------------------------------------
# workload construction (untimed)
# --------------------------------------------------------------------------------------------

def load_partition(train: torch.Tensor, num_nodes: int, shape: str) -> torch.Tensor:
    """METIS vector for the train graph: the committed fixture if it was computed on exactly these
    edges (checksum), otherwise METIS is run here (~75 s at ML-25M)."""
    fx = os.path.join(REPO, "tests", "golden", "ml25m_seed0_metis100.npz")
    if shape == "ml25m" and os.path.exists(fx):
        z = np.load(fx)
        if int(z["train_checksum"]) == int((train[0] * 31 + train[1]).sum()) and int(z["num_parts"]) == NUM_PARTS:
            return torch.from_numpy(z["cluster"].astype(np.int64))
    from lgcn_b200.data.dataset_handler import metis_partition
    return metis_partition(train, num_nodes, NUM_PARTS)


def cluster_batches_cpu(train: torch.Tensor, cluster: torch.Tensor, num_nodes: int):
    """Host-side list of the 100 cluster batches (global ids) for the CPU arms; same result as K4
    (tests/test_gpu_cluster_score.py) -- train is (row, col)-sorted so a stable selection suffices."""
    cr, cc = cluster[train[0]], cluster[train[1]]
    keep = cr == cc
    e, part = train[:, keep], cr[keep]
    order = torch.sort(part, stable=True)[1]
    e, part = e[:, order], part[order]
    cnt = torch.bincount(part, minlength=NUM_PARTS)
    off = torch.zeros(NUM_PARTS + 1, dtype=torch.long)
    off[1:] = torch.cumsum(cnt, 0)
    return [e[:, off[p]:off[p + 1]].contiguous() for p in range(NUM_PARTS)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_bf16_tflops():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["bf16_tflops"]) if os.path.exists(p) else 1590.0


def measured_peak_gbs():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores (also `--impl reference`)
# --------------------------------------------------------------------------------------------

def cpu_sample_batches(batches):
    """Three cluster batches at the 25/50/75-th size percentile: the per-batch CPU cost is dominated
    by full-table (N x 64) work, so the epoch estimate is 100 x their mean time."""
    sizes = np.array([b.shape[1] for b in batches])
    order = np.argsort(sizes)
    return [int(order[int(q * (len(order) - 1))]) for q in (0.25, 0.5, 0.75)]


def run_cpu_arm(g, batches, k, steps, warmup):
    from oracle import reference_path as ref
    torch.set_num_threads(os.cpu_count() or 1)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    st = ref.TrainState(u0, i0, k)
    pick = cpu_sample_batches(batches)
    gen = torch.Generator().manual_seed(0)
    total_edges = int(sum(b.shape[1] for b in batches))

    def one_step():
        t0 = time.perf_counter()
        for p in pick:
            ei = batches[p]
            neg = torch.randint(0, g.num_items, (int((ei[0] < g.num_users).sum()),), generator=gen)
            st.step(ei, neg)
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    times = [one_step() for _ in range(steps)]
    per_batch = float(np.mean(times)) / len(pick)
    epoch_s = per_batch * len(batches)
    sample = (f"{len(pick)} of {len(batches)} cluster batches per step (25/50/75th size percentile: "
              f"{[int(batches[p].shape[1]) for p in pick]} edges), epoch time = {len(batches)} x mean batch time")
    return {"value": total_edges / epoch_s, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": sample, "epoch_ms_estimate": epoch_s * 1e3, "ms_per_batch": per_batch * 1e3}


# --------------------------------------------------------------------------------------------
# ours, N = 1 (C2)
# --------------------------------------------------------------------------------------------

def run_c2(args):
    from lgcn_b200 import _lib
    from lgcn_b200.data.dataset_handler import ClusterData, ClusterLoader, Data
    from lgcn_b200.models.light_gcn import LightGCN
    from lgcn_b200.utils import train_test as tt

    shape = {"c2": "ml25m", "c1": "ml100k"}[args.workload]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    g = synthetic.make_graph(shape, seed=0)
    k = synthetic.SHAPES[shape][3]
    train = g.edges("train")
    n = g.num_nodes
    cluster = load_partition(train, n, shape)
    t0 = time.perf_counter()
    cd = ClusterData(Data(edge_index=train.to(dev), num_nodes=n), NUM_PARTS, cluster=cluster)
    torch.cuda.synchronize()
    extract_ms = (time.perf_counter() - t0) * 1e3
    parts = [d for d in cd.parts]
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    model = LightGCN(g.num_users, g.num_items, num_layers=k).to(dev)
    with torch.no_grad():
        model.user_embedding.weight.copy_(u0)
        model.item_embedding.weight.copy_(i0)
    opt = tt.FusedAdam(model)
    loader = ClusterLoader(parts, shuffle=True)
    t0 = time.perf_counter()
    graphs = [model.graph(d.edge_index) for d in parts]          # K0 once per batch tensor (cached)
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3
    live = [d for d, gr in zip(parts, graphs) if gr.num_triplets > 0]
    edges_per_epoch = int(sum(d.edge_index.shape[1] for d in live))
    torch.manual_seed(0)

    def epoch():
        return tt.train(model, opt, loader, dev)

    for _ in range(args.warmup):
        epoch()
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    losses = [epoch() for _ in range(args.steps)]
    ev1.record()
    torch.cuda.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    ms_per_step = total_ms / args.steps
    value = edges_per_epoch / (ms_per_step * 1e-3)

    # ---- e2e: the same epoch through train() with HOST batches (pinned), H2D every step ----------
    host_parts = [Data(edge_index=d.edge_index.cpu().pin_memory(), num_nodes=n) for d in live]

    class HostLoader:
        """What the reference's DataLoader hands to train(): batches that live on the HOST (pinned);
        train() uploads them (utils/train_test.py:87 `batch.to(device)`) every epoch."""
        def __iter__(self):
            return iter(host_parts)

    model._graphs.capacity = 4                                     # per-step uploads must not pile up
    for _ in range(2):
        tt.train(model, opt, HostLoader(), dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(1, args.steps)
    e0.record()
    for _ in range(e2e_steps):
        tt.train(model, opt, HostLoader(), dev)                    # returns after the loss D2H
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    model._graphs.capacity = 512
    h2d = int(sum(hp.edge_index.numel() * 8 for hp in host_parts))
    e2e = {"value": edges_per_epoch / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * len(host_parts) + 8 * (16 * len(host_parts) + 1),
           "note": "train() over HOST (pinned) batches: every epoch uploads all edge lists, rebuilds every batch's "
                   "normalisation/CSR (one batched K0b call) and reads the losses back"}

    # ---- per-stage device times of the DENSE step over one epoch (events on the launch stream) ----
    stage_ms = stage_breakdown(model, opt, live, graphs_of(model, live), k, dev)
    peak, peak_src = measured_peak_gbs()
    adam_bytes = 7 * n * 256                                       # p,m,v,grad read + p,m,v written
    adam_ms = stage_ms["clip_adam"] / len(live)
    spmm = full_graph_propagation(model, train.to(dev), k, dev, peak)
    # roofline kernel = the propagation layer (SURVEY.md sec.8d defines B_layer for it): one launch = one
    # layer over the train graph, duration = event time of the K-layer call / K
    b_layer = 2 * n * 256 + 8 * train.shape[1] + 4 * (n + 1)
    layer_ms = spmm["ms"] / k
    roofline = {"kernel": "rowtask_kernel<FwdOp> (one propagation layer, full train graph)", "bound": "hbm",
                "achieved": b_layer / (layer_ms * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "traffic": ncu_traffic("rowtask_fwd_layer"), "algorithmic_bytes_per_launch": b_layer,
                "avg_launch_ms": layer_ms,
                "note": "table (56.7 MB) is L2-resident, so the kernel is bound by the L2 gather rate, not by the compulsory "
                        "HBM bytes: see l2_gather (gather-model bytes E*264 + N*256 per layer against the measured "
                        "random-row-gather ceiling of the L2)"}
    roofline["frac"] = roofline["achieved"] / peak
    roofline["l2_gather"] = {"achieved": spmm["gather_model_gbs"], "peak": spmm["l2_gather_ceiling_gbs"], "unit": "GB/s",
                             "frac": spmm["frac_of_l2_gather_ceiling"], "peak_source": spmm["l2_gather_ceiling_how"]}
    roofline_adam = {"kernel": "clip_adam_kernel (dense step)", "bound": "hbm",
                     "achieved": adam_bytes / (adam_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "traffic": ncu_traffic("clip_adam"), "algorithmic_bytes_per_launch": adam_bytes,
                     "avg_launch_ms": adam_ms}
    roofline_adam["frac"] = roofline_adam["achieved"] / peak
    # the kernel the epoch actually spends its time in: one persistent launch per epoch (epoch_kernel.cu).
    # Algorithmic bytes per step: Adam state of the touched rows (p, m, v read + written) + the activation rows
    # of the active nodes (y0..y_{K-1}, final, G, z tables, grad: written once, read once or K times) + indices.
    grs = graphs_of(model, live)
    step_bytes = 0
    for gr in grs:
        touched = gr.num_active + min(gr.num_triplets, g.num_items)
        step_bytes += touched * 256 * 6 + gr.num_active * 256 * (2 * (2 * k + 3)) + gr.num_edges * 4 * (2 * k + 4)
    roofline_step = {"kernel": "epoch_kernel (lgcn_train_steps_sparse: all sparse steps of an epoch in one launch)",
                     "bound": "hbm", "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "algorithmic_bytes_per_launch": step_bytes, "avg_launch_ms": ms_per_step,
                     "traffic": ncu_traffic("epoch_kernel"),
                     "note": "latency-bound by construction: ~2 k rows and ~7 k edges per step, 2K+2 device-wide barriers "
                             "per step (see profiles/r1c_epoch_trace.txt for where a step goes); duration here is the whole "
                             "epoch (kernel + end-of-epoch flush)"}
    roofline_step["frac"] = roofline_step["achieved"] / peak
    n_sparse = sum(1 for gr in grs if tt.sparse_step_pays(gr))
    # per epoch: the sparse batches run inside ONE persistent cooperative launch (epoch_kernel) followed by one
    # adam_replay_kernel (flush); a dense batch is launches_per_step(k, False) kernels
    launches = args.steps * ((2 if n_sparse else 0) + (len(live) - n_sparse) * launches_per_step(k, False))

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"C2 Cluster-GCN training epoch, {shape} shape: U={g.num_users} I={g.num_items} "
                                  f"E_train={train.shape[1]} directed, {NUM_PARTS} METIS parts "
                                  f"({len(live)} non-empty, {edges_per_epoch} intra-cluster edges), K={k}, dim=64, "
                                  "fwd+BPR+bwd+clip+Adam per batch, step = 1 epoch",
                      "l2": "an epoch touches ~0.9 GB of distinct rows/state (100 batches x ~9 MB) plus 24 MB of CSR, "
                            "more than the 126 MB L2; no explicit flush",
                      "parallelism": "1 GPU",
                      "step_kinds": f"{n_sparse} touched-rows (sparse) steps inside one persistent cooperative launch "
                                    f"(lgcn_train_steps_sparse) + {len(live) - n_sparse} dense steps per epoch"},
           "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
           "roofline": roofline, "roofline_adam": roofline_adam, "roofline_step_kernel": roofline_step,
           "dense_stage_ms_per_epoch": stage_ms,
           "spmm_full_graph": spmm,
           "setup_ms": {"cluster_extract": extract_ms, "graph_build_100_batches": build_ms},
           "final_epoch_loss": losses[-1]}
    if args.workload == "c2" and not args.no_eval:
        try:
            out["full_graph_step_1gpu_c3"] = full_graph_step_single_gpu(g, k, dev)
        except Exception as exc:                                    # an extra, never the reason for a missing line
            out["full_graph_step_1gpu_c3"] = {"error": repr(exc)[:200]}
    if not args.no_cpu:
        out["cpu_baseline"] = run_cpu_arm(g, cluster_batches_cpu(train, cluster, n), k, 2, 1)
    return out


def full_graph_step_single_gpu(g, k, dev, steps=10, warmup=3):
    """The N > 1 workload (C3: one full-graph training step, node-range sharded) on ONE GPU, so that the per-N lines
    of the scaling run can be set against a single-GPU run of the SAME workload (the N = 1 line itself is C2)."""
    from lgcn_b200 import sharded
    tr = g.edges("train").to(dev)
    ops = sharded.CudaOps(tr, g.num_users, g.num_items, k)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    trainer = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
    for _ in range(warmup):
        trainer.step_sampled(g.num_items, use_graph=False)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        trainer.step_sampled(g.num_items, use_graph=False)
    z.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(z) / steps
    return {"workload": "C3 full-graph training step on 1 GPU (what `--gpus N` runs sharded)", "ms_per_step": ms,
            "edges_per_s": tr.shape[1] / (ms * 1e-3), "steps": steps, "warmup": warmup}


def launches_per_step(k: int, sparse: bool) -> int:
    if sparse:   # step_begin, mark_negs, 2 replays, K fwd, BPR A+B, K bwd, neg_rows_grad, 2 adam_rows
        return 1 + 1 + 2 + k + 2 + k + 1 + 2
    # step_begin + inactive-row fwd + K fwd layers + BPR pass A + pass B + inactive-row bwd + K bwd + clip_adam
    return 1 + 1 + k + 2 + 1 + k + 1


def ncu_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the last committed ncu --set full
    capture (profiles/roofline_traffic.json), or None."""
    p = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(key, {}).get("dram_bytes")


def graphs_of(model, parts):
    return [model.graph(d.edge_index) for d in parts]


def stage_breakdown(model, opt, parts, graphs, k, dev):
    """One epoch through the fine-grained C-ABI calls with an event pair around every stage."""
    from ctypes import byref
    from lgcn_b200 import _lib
    L = _lib.lib()
    s = _lib.stream_ptr(dev)
    b = opt.buffers
    uw, iw = model.user_embedding.weight, model.item_embedding.weight
    names = ["step_begin", "propagate_fwd", "bpr_fwd_bwd", "propagate_bwd", "clip_adam"]
    evs = {nm: [] for nm in names}
    loss = torch.empty(1, device=dev)
    for d, g in zip(parts, graphs):
        neg = torch.randint(0, model.num_items, (g.num_triplets,), device=dev)
        b.ensure_triplets(g.num_triplets)
        reg = 2.0 * 5e-3 / (64.0 * g.num_triplets)
        calls = [
            lambda: L.lgcn_step_begin(byref(opt.c), b.accum.data_ptr(), s),
            lambda: L.lgcn_propagate_fwd(g.ref, uw.data_ptr(), iw.data_ptr(), k, b.final_emb.data_ptr(),
                                         b.rnorm.data_ptr(), b.work.data_ptr(), b.work.numel() * 4, s),
            lambda: L.lgcn_bpr_fwd_bwd(g.ref, b.final_emb.data_ptr(), b.rnorm.data_ptr(), neg.data_ptr(),
                                       b.grad_final.data_ptr(), b.neg_count.data_ptr(), b.trip_scratch.data_ptr(),
                                       b.accum.data_ptr(), s),
            lambda: L.lgcn_propagate_bwd(g.ref, b.grad_final.data_ptr(), k, uw.data_ptr(), iw.data_ptr(),
                                         b.neg_count.data_ptr(), reg, b.grad_e0.data_ptr(), b.accum.data_ptr(),
                                         b.work.data_ptr(), b.work.numel() * 4, s),
            lambda: L.lgcn_clip_adam(byref(opt.c), uw.data_ptr(), iw.data_ptr(), model.num_users, model.num_items,
                                     b.grad_e0.data_ptr(), b.accum.data_ptr(), g.num_triplets, 5e-3, loss.data_ptr(), s),
        ]
        for nm, fn in zip(names, calls):
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(fn())
            z.record()
            evs[nm].append((a, z))
    torch.cuda.synchronize()
    return {nm: float(sum(a.elapsed_time(z) for a, z in v)) for nm, v in evs.items()}


def full_graph_propagation(model, train_dev, k, dev, peak):
    """The SpMM figures of the metric: K-layer fused propagation over the whole train graph."""
    from lgcn_b200 import _lib
    L = _lib.lib()
    g = model.graph(train_dev)
    n, e = g.num_nodes, g.num_edges
    final = torch.empty(n, 64, device=dev)
    work = torch.empty(max(k - 1, 1) * n * 64, device=dev)
    uw, iw = model.user_embedding.weight, model.item_embedding.weight

    def run():
        _lib.check(L.lgcn_propagate_fwd(g.ref, uw.data_ptr(), iw.data_ptr(), k, final.data_ptr(), None,
                                        work.data_ptr(), work.numel() * 4, _lib.stream_ptr(dev)))
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); z.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(z))
    ms = float(np.median(ts))
    b_layer = 2 * n * 256 + 8 * e + 4 * (n + 1)                    # SURVEY.md sec.8(d) contract figure
    gather = e * 264 + n * 256 + 4 * (n + 1)
    model._graphs.clear()
    # measured ceiling of this access shape: independent random 256-byte row gathers over a table of the same
    # size, no index array, no dependent address (csrc/probe.cu) -- what the L2 can deliver at best
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    table = torch.randn(n, 64, device=dev)
    sink = torch.zeros(sms * 6 * 16, device=dev)
    rows_hw = 4096

    def probe():
        _lib.check(L.lgcn_probe_gather(table.data_ptr(), n, sms * 6, rows_hw, sink.data_ptr(), _lib.stream_ptr(dev)))
    for _ in range(3):
        probe()
    pt = []
    for _ in range(5):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); probe(); z.record()
        torch.cuda.synchronize()
        pt.append(a.elapsed_time(z))
    l2_peak = sms * 6 * 16 * rows_hw * 256 / (float(np.median(pt)) * 1e-3) / 1e9
    gather_gbs = k * gather / (ms * 1e-3) / 1e9
    return {"edges": e, "layers": k, "ms": ms, "edges_per_s": e * k / (ms * 1e-3),
            "algorithmic_gbs": k * b_layer / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": k * b_layer / (ms * 1e-3) / 1e9 / peak,
            "gather_model_gbs": gather_gbs, "l2_gather_ceiling_gbs": l2_peak, "frac_of_l2_gather_ceiling": gather_gbs / l2_peak,
            "l2_gather_ceiling_how": "lgcn_probe_gather: random 256 B row gathers over an equally sized (L2-resident) table, "
                                     "measured live"}


# --------------------------------------------------------------------------------------------
# C4: full-rank evaluation (reported inside the C2 line and on its own with --workload c4)
# --------------------------------------------------------------------------------------------

def run_c4(args, shape="ml25m"):
    from lgcn_b200.utils import recommend as rec
    dev = torch.device("cuda:0")
    g = synthetic.make_graph(shape, seed=0)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    ue, ie = u0.to(dev), i0.to(dev)
    train, test = g.edges("train").to(dev), g.edges("test").to(dev)
    ptr, idx = rec.exclusion_csr(train, g.num_users)
    k = 20

    def timed(algo):
        def run():
            return rec.score_topk(ue, ie, k, True, ptr, idx, algo=algo)
        for _ in range(2):
            run()
        ts = []
        for _ in range(max(3, min(args.steps, 10))):
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); z.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(z))
        return float(np.median(ts))
    ms = timed(rec.SCORE_TENSOR)
    ms_ffma = timed(rec.SCORE_FFMA)
    m = rec.full_rank_eval(ue, ie, train, test, g.num_users, k=k)
    flop = 2.0 * g.num_users * g.num_items * 64
    return {"workload": f"C4 full-rank eval {g.num_users} x {g.num_items} x 64, train-edge mask, top-{k}",
            "ms": ms, "scores_per_s": g.num_users * g.num_items / (ms * 1e-3),
            "useful_tflops": flop / (ms * 1e-3) / 1e12, "issued_tf32_tflops": 3 * flop / (ms * 1e-3) / 1e12,
            "tensor_frac_of_tf32_peak": 3 * flop / (ms * 1e-3) / 1e12 / (measured_bf16_tflops() / 2),
            "math": "tcgen05.mma kind::tf32, 3-term hi/lo split (fp32-level accuracy), fp32 accumulate in TMEM",
            "ffma_kernel_ms": ms_ffma, "recall@20": m["recall"], "ndcg@20": m["ndcg"],
            "users_with_test_items": m["users"], "note": "random-init embeddings: recall/NDCG are chance level"}


# --------------------------------------------------------------------------------------------
# ours, N >= 1 under torchrun (C3): node-range sharded full-graph training step
# --------------------------------------------------------------------------------------------

def run_c3(args):
    os.environ["NCCL_DEBUG"] = os.environ.get("LGCN_NCCL_DEBUG", "WARN")      # keep NCCL's banner off stdout
    import torch.distributed as dist
    from lgcn_b200 import sharded
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                      # NCCL prints its version banner to stdout on first use
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    shape = os.environ.get("LGCN_BENCH_SHAPE", "ml25m")
    g = synthetic.make_graph(shape, seed=0)
    k = synthetic.SHAPES[shape][3]
    train = g.edges("train")
    n, e = g.num_nodes, train.shape[1]
    tr = train.to(dev)
    ops = sharded.CudaOps(tr, g.num_users, g.num_items, k)
    u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
    trainer = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
    exchange = ("single GPU, no exchange" if world == 1 else
                ("fused into the SpMM epilogue: " + ("NVLS multicast stores" if ops.multicast else "per-peer NVLink stores")
                 + " into symmetric memory + barrier per layer") if ops.p2p else
                f"NCCL all-gather between kernels (symmetric memory unavailable: {ops.p2p_error})")
    p = ops.num_triplets
    torch.manual_seed(0)                       # same Philox stream on every rank => identical negatives

    use_graph = os.environ.get("LGCN_SHARDED_GRAPH", "0") == "1"     # opt-in until validated at every N

    def step():
        return trainer.step_sampled(g.num_items, use_graph=use_graph)

    for _ in range(max(args.warmup, 5 if use_graph else 0)):      # 3 eager + capture + 1 replay
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    trainer.comm.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss = step()
    ev1.record()
    torch.cuda.synchronize()
    trainer.comm.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / args.steps
    clk = clocks.stop() if rank == 0 else None

    # forward propagation alone: per-layer SpMM + all-gather figures of BASELINE config C3
    for _ in range(2):
        trainer.propagate_only()
    torch.cuda.synchronize(); trainer.comm.barrier()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        trainer.propagate_only()
    z.record()
    torch.cuda.synchronize()
    tp = torch.tensor([a.elapsed_time(z) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    prop_ms = float(tp)

    # e2e: edge list uploaded from pinned host memory, CSR rebuilt, one step, loss read back
    host_ei = train.pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    from lgcn_b200 import _lib

    def e2e_step():
        ei = host_ei.to(dev, non_blocking=True)
        fresh = _lib.Graph(ei, g.num_users, g.num_items)           # K0 on the uploaded edge list
        del fresh                                                   # (same content: the trainer keeps its CSR)
        return float(step().item())

    for _ in range(2):                                              # untimed: the caching allocator gets its blocks
        e2e_step()
    torch.cuda.synchronize(); trainer.comm.barrier()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(e2e_steps):
        e2e_step()
    z.record()
    torch.cuda.synchronize()
    te = torch.tensor([a.elapsed_time(z) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    peak, peak_src = measured_peak_gbs()
    b_layer = 2 * n * 256 + 8 * e + 4 * (n + 1)
    roof = {"kernel": "rowtask_kernel<FwdOp> (K-layer propagation incl. all-gathers)", "bound": "hbm",
            "achieved": k * b_layer / (prop_ms * 1e-3) / 1e9, "peak": peak * world, "peak_source": peak_src + f" x {world} GPUs",
            "unit": "GB/s", "traffic": None, "algorithmic_bytes_per_launch": k * b_layer, "avg_launch_ms": prop_ms}
    roof["frac"] = roof["achieved"] / roof["peak"]
    return {"metric": METRIC, "value": e / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C3 full-graph training step, {shape} shape: U={g.num_users} I={g.num_items} "
                                   f"E_train={e} directed, K={k}, dim=64, fwd+BPR+bwd+clip+Adam, node-range sharded "
                                   f"over {world} GPU(s) with all-gather per layer + all-reduce of dL/dfinal",
                       "l2": "tables + activations + CSR (> 1 GB) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"node-range x{world}", "exchange": exchange,
                       "launch": ("one CUDA graph per step (sampling + kernels + barriers + NCCL)"
                                  if getattr(trainer, "_graph", None) is not None else
                                  "eager launches" + (f" (graph capture failed: {trainer._graph_error})"
                                                      if getattr(trainer, "_graph_failed", False) else ""))},
            "clocks": clk, "gpu_launches": int(args.steps * (2 + 2 + 2 * k + 2 + 2 * k + 2 + 1)),
            "e2e": {"value": e / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(train.numel() * 8), "d2h_bytes_per_step": 4,
                    "note": "edge list uploaded from pinned host memory and CSR rebuilt every step (the reference "
                            "re-derives the normalisation from edge_index in every forward)"},
            "roofline": roof, "propagation": {"ms": prop_ms, "edges_per_s": e * k / (prop_ms * 1e-3), "layers": k},
            "final_loss": float(loss)}


# --------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------

def run_reference(args):
    """The reference's CPU implementation of the path = the oracle port (PyG / torch_sparse are not
    installable offline and /root/reference does not exist on the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    shape = {"c2": "ml25m", "c1": "ml100k", "c3": os.environ.get("LGCN_BENCH_SHAPE", "ml25m")}[args.workload]
    g = synthetic.make_graph(shape, seed=0)
    k = synthetic.SHAPES[shape][3]
    train = g.edges("train")
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if args.workload == "c3":
        from oracle import reference_path as ref
        torch.set_num_threads(os.cpu_count() or 1)
        stride = 16
        sub = train[:, ::stride].contiguous()
        u0, i0 = synthetic.init_embeddings(g.num_users, g.num_items, 64, 0)
        st = ref.TrainState(u0, i0, k)
        gen = torch.Generator().manual_seed(0)
        p = int((sub[0] < g.num_users).sum())

        def one():
            t0 = time.perf_counter()
            st.step(sub, torch.randint(0, g.num_items, (p,), generator=gen))
            return time.perf_counter() - t0
        for _ in range(min(args.warmup, 1)):
            one()
        ts = [one() for _ in range(min(args.steps, 5))]
        sec = float(np.mean(ts))
        cpu = {"value": sub.shape[1] / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"full-table training step on every {stride}th train edge ({sub.shape[1]} of {train.shape[1]} "
                         f"edges, N unchanged); edges/s of the sample, {len(ts)} timed steps"}
        base.update({"value": cpu["value"], "ms_per_step": sec * 1e3 * stride, "scaling": "strong",
                     "config": {"workload": f"C3 full-graph training step, {shape} shape, K={k}, dim=64; reference CPU path "
                                            "= oracle port of the PyG gather/scatter op sequence, torch CPU, all host threads"},
                     "cpu_baseline": cpu,
                     "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return base
    cluster = load_partition(train, g.num_nodes, shape)
    batches = [b for b in cluster_batches_cpu(train, cluster, g.num_nodes) if int((b[0] < g.num_users).sum()) > 0]
    cpu = run_cpu_arm(g, batches, k, args.steps, args.warmup)
    base.update({"value": cpu["value"], "ms_per_step": cpu["epoch_ms_estimate"], "scaling": "weak",
                 "config": {"workload": f"C2 Cluster-GCN training epoch, {shape} shape, {NUM_PARTS} METIS parts, K={k}, dim=64; "
                                        "reference CPU path = oracle port of the PyG gather/scatter op sequence "
                                        "(PyG/torch_sparse not installable offline), torch CPU, all host threads"},
                 "cpu_baseline": cpu,
                 "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eval", action="store_true", help="skip the C4 full-rank evaluation section")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "c2" if args.gpus == 1 else "c3"
    if args.impl == "reference":
        out = run_reference(args)
        if out is not None:
            print(json.dumps(out))
        return
    if args.workload == "c4":
        out = run_c4(args)
    elif args.gpus == 1 and args.workload in ("c1", "c2"):
        out = run_c2(args)
        if args.workload == "c2" and not args.no_eval:
            out["eval_full_rank_c4"] = run_c4(args)
    else:
        out = run_c3(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
