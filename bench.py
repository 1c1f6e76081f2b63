#!/usr/bin/env python
"""bench.py -- LightGCN hot-path benchmark (contract in the task statement / DESIGN.md sec. Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c5|c2|c4]

EVERY N runs the SAME workload, so the per-N lines are a like-for-like strong-scaling curve:

  C3 (BASELINE.json configs[2], default)  one full-graph LightGCN training step -- forward (K = 3 propagation
      layers), BPR loss, backward, clip, Adam (utils/train_test.py:88-96) -- on the synthetic ML-25M-shaped
      graph (162,541 users x 59,047 movies, 25.0 M directed edges, 22.5 M of them train), node-range sharded
      over the N GPUs (lgcn_b200/sharded.py).  A full-graph step IS a full-graph epoch.
  C5 (--workload c5, configs[4])          the same step on the 10x graph (1.6 M x 0.6 M, 225 M train edges, K = 4).

metric  lightgcn_train_edges_per_s = directed train edges consumed per second by full training steps
        (whole job); ms_per_step is the step (= full-graph epoch) time, max over ranks.
At N = 1 (C3) the line also carries, as named blocks with their own roofline / e2e / cpu_baseline:
  cluster_gcn_epoch_c2   BASELINE configs[1]: one Cluster-GCN epoch (100 METIS parts) through train()
  eval_full_rank_c4      BASELINE configs[3]: 162,541 x 59,047 scoring + train mask + top-20
  spmm_full_graph        K-layer propagation alone (edges/s, GB/s against the HBM and the L2-gather ceilings)
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
SHAPE_OF = {"c3": "ml25m", "c5": "ml25m_x10", "c2": "ml25m", "c1": "ml100k", "c4": "ml25m"}


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


def ncu_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the last committed ncu --set full
    capture (profiles/roofline_traffic.json), or None."""
    p = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(key, {}).get("dram_bytes")


_REAL_STDOUT = None


def emit_line(obj) -> None:
    """The ONE JSON line on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, line)


def event_pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


# --------------------------------------------------------------------------------------------
# CPU arms: the oracle port timed on the host cores (cpu_baseline blocks and `--impl reference`)
# --------------------------------------------------------------------------------------------

def cpu_full_graph_sample(nu, ni, train, k, steps, warmup, budget_s):
    """The reference's CPU implementation of a full-graph training step (oracle port of the PyG gather / scatter op
    sequence) on a BOUNDED sample of the workload: every `stride`-th train edge, node set unchanged, so that
    (steps + warmup) sample steps fit `budget_s`.  Returns the cpu_baseline dict + the per-step seconds."""
    from oracle import reference_path as ref
    torch.set_num_threads(os.cpu_count() or 1)
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    gen = torch.Generator().manual_seed(0)

    def run(stride, reps):
        sub = train[:, ::stride].contiguous()
        st = ref.TrainState(u0, i0, k)
        p = int((sub[0] < nu).sum())
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            st.step(sub, torch.randint(0, ni, (p,), generator=gen))
            ts.append(time.perf_counter() - t0)
        return sub.shape[1], ts

    _, probe = run(32, 2)                                  # ~1-2 s: sizes the sample
    per_edge_32 = probe[-1]
    stride = 32
    for cand in (16, 8, 4, 2, 1):                          # cost grows roughly linearly with the edge count
        if per_edge_32 * (32 / cand) * (steps + warmup) <= budget_s:
            stride = cand
    edges, ts = run(stride, steps + warmup)
    ts = ts[warmup:]
    sec = float(np.mean(ts))
    return {"value": edges / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"full-table training step on every {stride}th train edge ({edges} of {train.shape[1]} edges, "
                      f"all {nu + ni} nodes); edges/s of the sample, {len(ts)} timed steps of {sec:.2f} s",
            "ms_per_sample_step": sec * 1e3, "stride": stride,
            "full_step_ms_estimate": sec * 1e3 * train.shape[1] / edges}, sec


def cpu_cluster_epoch(nu, ni, batches, k):
    """One WHOLE Cluster-GCN epoch on the CPU oracle: all batches, the hub cluster included, once."""
    from oracle import reference_path as ref
    torch.set_num_threads(os.cpu_count() or 1)
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    st = ref.TrainState(u0, i0, k)
    gen = torch.Generator().manual_seed(0)
    total_edges = int(sum(b.shape[1] for b in batches))
    st.step(batches[1], torch.randint(0, ni, (int((batches[1][0] < nu).sum()),), generator=gen))   # warm-up
    t0 = time.perf_counter()
    for ei in batches:
        st.step(ei, torch.randint(0, ni, (int((ei[0] < nu).sum()),), generator=gen))
    sec = time.perf_counter() - t0
    return {"value": total_edges / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"the whole epoch once: all {len(batches)} cluster batches incl. the hub cluster "
                      f"({max(b.shape[1] for b in batches)} edges), {sec:.1f} s", "epoch_ms": sec * 1e3}


# --------------------------------------------------------------------------------------------
# ours: the full-graph training step (C3 / C5), any N
# --------------------------------------------------------------------------------------------

def staged_step(trainer, neg, acc):
    """One eager step with an event pair around every stage (same sequence as ShardedTrainer.step)."""
    o, k = trainer.ops, trainer.k
    p2p = getattr(o, "p2p", False)

    def timed(name, fn):
        a, z = event_pair()
        a.record(); fn(); z.record()
        acc.setdefault(name, []).append((a, z))
    timed("step_begin+prescale", lambda: (o.step_begin(), o.prescale()))
    timed("bpr_buckets", lambda: o.bpr_buckets(neg))               # overlaps the y_0 exchange (see ShardedTrainer.step)
    timed("exchange", lambda: trainer._gather(o.y[0]))
    for layer in range(1, k + 1):
        timed("spmm_fwd_layer", lambda layer=layer: o.fwd_layer(layer))
        timed("exchange", lambda layer=layer: trainer._gather(o.y[layer] if layer < k else o.final))
    timed("bpr", lambda: o.bpr(neg))
    timed("exchange", lambda: trainer._gather(o.zg))
    for j in range(1, k + 1):
        timed("spmm_bwd_layer", lambda j=j: o.bwd_layer(j, trainer.bpr_coeff))
        if j < k:
            timed("exchange", lambda j=j: trainer._gather(o.zbuf(j)))
    timed("exchange", lambda: o.allreduce_accum() if p2p else trainer.comm.allreduce(o.accum))
    timed("clip_adam", lambda: o.clip_adam(trainer.bpr_coeff))


def run_full_graph(args):
    import torch.distributed as dist
    from lgcn_b200 import sharded
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        # NCCL logs to stdout (INIT lines name the communicator's nranks): for the rest of the run fd 1 is stderr, and
        # the JSON line goes out through the saved descriptor (emit_line), so stdout carries nothing else
        if os.environ.get("NCCL_DEBUG", "WARN").upper() in ("WARN", "VERSION"):     # the INIT lines are the evidence of nranks / NVLS
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        sys.stdout.flush()
        global _REAL_STDOUT
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    shape = SHAPE_OF[args.workload]
    k = synthetic.SHAPES[shape][3]
    # the 10x graph is drawn with the CUDA generator (seconds instead of ~90 s on the host): same recipe, another
    # seeded graph than the CPU-generated one of the fixtures; its checker (--check) runs on this very edge list
    gen_dev = dev if shape == "ml25m_x10" else None
    nu, ni, train, graph = synthetic.shared_train_edges(shape, local, (dist.barrier if world > 1 else (lambda: None)),
                                                        device=gen_dev)
    n, e = nu + ni, train.shape[1]
    ops = sharded.CudaOps(train.to(dev), nu, ni, k, device=dev)      # the shard is cut on the device
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    trainer = sharded.ShardedTrainer(ops, u0.to(dev), i0.to(dev), sharded.Comm())
    exchange = ("single GPU, no exchange" if world == 1 else
                ("fused into the kernel epilogues: " + ("NVLS multicast stores" if ops.multicast else "per-peer NVLink stores")
                 + " into symmetric memory + one peer barrier per table; no NCCL call inside the step") if ops.p2p else
                f"NCCL all-gather between kernels (symmetric memory unavailable: {ops.p2p_error})")
    torch.manual_seed(0)                       # same Philox stream on every rank => identical negatives
    use_graph = os.environ.get("LGCN_SHARDED_GRAPH", "1") == "1"

    def step():
        return trainer.step_sampled(ni, use_graph=use_graph)

    for _ in range(max(args.warmup, 5 if use_graph else 0)):      # 3 eager + capture + 1 replay
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    trainer.comm.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = event_pair()
    ev0.record()
    for _ in range(args.steps):
        loss = step()
    ev1.record()
    torch.cuda.synchronize()
    trainer.comm.barrier()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    clk = clocks.stop() if rank == 0 else None
    final_loss = float(loss)

    # ---- per-stage device times: an instrumented EAGER pass right after the timed region (events cannot be placed
    # inside a CUDA-graph replay); per stage the max over ranks of the mean over `inst` steps
    inst = max(3, min(args.steps, 10))
    acc = {}
    neg = torch.randint(0, ni, (ops.num_triplets,), device=dev)
    staged_step(trainer, neg, {})
    torch.cuda.synchronize(); trainer.comm.barrier()
    for _ in range(inst):
        staged_step(trainer, neg, acc)
    torch.cuda.synchronize()
    stage_ms = {nm: max_over_ranks(sum(a.elapsed_time(z) for a, z in v) / inst) for nm, v in sorted(acc.items())}
    n_layers = {nm: len(v) // inst for nm, v in acc.items()}
    layer_ms = (stage_ms["spmm_fwd_layer"] + stage_ms["spmm_bwd_layer"]) / (2 * k)

    # ---- forward propagation alone (BASELINE "propagation edges/s")
    for _ in range(2):
        trainer.propagate_only()
    torch.cuda.synchronize(); trainer.comm.barrier()
    a, z = event_pair()
    a.record()
    for _ in range(5):
        trainer.propagate_only()
    z.record()
    torch.cuda.synchronize()
    prop_ms = max_over_ranks(a.elapsed_time(z) / 5)

    # ---- e2e: every step uploads this rank's shard of the edge list from PINNED host memory, rebuilds its CSR pair
    # (K0), re-derives the triplet index, runs one step and reads the loss back
    sh = ops.shard
    host_edges = sh.edges.cpu().to(torch.int64).pin_memory()
    host_trip = None if sh.trip_global is None else sh.trip_global.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    trainer.drop_graph()

    def e2e_step_serial():
        ed = host_edges.to(dev, non_blocking=True)
        tg = None if host_trip is None else host_trip.to(dev, non_blocking=True)
        ops.load_shard(ed, tg)
        return float(trainer.step_sampled(ni, use_graph=False).item())

    def timed_e2e(fn, reps):
        for _ in range(2):                                          # untimed: the caching allocator gets its blocks
            fn()
        torch.cuda.synchronize(); trainer.comm.barrier()
        a, z = event_pair()
        a.record()
        for _ in range(reps):
            fn()
        z.record()
        torch.cuda.synchronize()
        return max_over_ranks(a.elapsed_time(z) / reps)
    e2e_serial_ms = timed_e2e(e2e_step_serial, 3)
    pipe = sharded.HostShardPipeline(trainer, host_edges, host_trip)
    e2e_ms = timed_e2e(lambda: float(pipe.step(ni).item()), e2e_steps)
    h2d = int(host_edges.numel() * 8 + (0 if host_trip is None else host_trip.numel() * 4))
    h2d_all = torch.tensor([h2d], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(h2d_all)

    peak, peak_src = measured_peak_gbs()
    b_layer = 2 * n * 256 + 8 * e + 4 * (n + 1)                     # SURVEY.md sec.8(d) contract figure, whole graph
    roof = {"kernel": "rowtask_kernel<FwdOp|BwdOp> (one propagation layer = one launch per rank; 2K launches per step)",
            "bound": "hbm", "achieved": b_layer / (layer_ms * 1e-3) / 1e9, "peak": peak * world,
            "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""), "unit": "GB/s",
            "traffic": ncu_traffic("rowtask_fwd_layer_r2") if world == 1 and shape == "ml25m" else None,
            "algorithmic_bytes_per_launch": b_layer, "avg_launch_ms": layer_ms,
            "share_of_step": 2 * k * layer_ms / sum(stage_ms.values()),
            "how": f"mean of the {2 * k} layer launches per step over {inst} instrumented eager steps right after the "
                   "timed region (CUDA events on the launch stream), max over ranks; bytes = compulsory bytes of one "
                   "layer over the WHOLE graph (all ranks together)",
            "note": ("the gathered table is L2-resident at this shape, so the layer is bound by the L2 gather rate, not by "
                     "compulsory HBM bytes: see spmm_full_graph.frac_of_l2_gather_ceiling") if shape == "ml25m" else
                    "the gathered table (563 MB) does not fit the L2: gathers stream from HBM"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    launches_per_step = 1 + 2 + k + 7 + k + 2 + len(ops.inactive_segs) * 2 + (2 * k + 3 if world > 1 else 0)
    out = {"metric": METRIC, "value": e / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32",
           "data": "synthetic" + (" (CUDA-generated seeded graph)" if gen_dev is not None else ""),
           "config": {"workload": f"{args.workload.upper()} full-graph training step (= full-graph epoch), {shape} shape: "
                                  f"U={nu} I={ni} E_train={e} directed, K={k}, dim=64, fwd+BPR+bwd+clip+Adam, node-range "
                                  f"sharded over N GPU(s)",
                      "l2": "tables + activations + CSR (> 1 GB) exceed the 126 MB L2; no explicit flush",
                      "parallelism": f"node-range x{world}", "exchange": exchange,
                      "shard_edges_rank0": int(ops.g.num_edges),
                      "launch": ("one CUDA graph per step (sampling + kernels + peer barriers)"
                                 if use_graph and not getattr(trainer, "_graph_failed", False) else
                                 "eager launches" + (f" (graph capture failed: {trainer._graph_error})"
                                                     if getattr(trainer, "_graph_failed", False) else ""))},
           "clocks": clk, "gpu_launches": int(args.steps * launches_per_step),
           "e2e": {"value": e / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": 4 * world,
                   "serial_ms_per_step": e2e_serial_ms,
                   "note": "per step every rank uploads ITS SHARD of the edge list (pinned int64 [2,E_r] + global triplet "
                           "numbers) and rebuilds its CSR pair / normalisation (the reference re-derives the normalisation "
                           "from edge_index in every forward), runs the step eagerly and reads the loss back (a host sync "
                           "per step); sharded.HostShardPipeline: the upload of step i+1 runs on a copy stream while step "
                           "i builds and trains (two device buffers), every step still uploads and rebuilds its own "
                           "graph; serial_ms_per_step = the same without the overlap; h2d bytes are summed over ranks"},
           "roofline": roof, "stage_ms_per_step": stage_ms, "stage_launches_per_step": n_layers,
           "propagation": {"ms": prop_ms, "edges_per_s": e * k / (prop_ms * 1e-3), "layers": k},
           "final_loss": final_loss}
    if world > 1:
        out["exchange_bytes"] = {"per_table_per_gpu_received": int(n * 256 * (world - 1) / world),
                                 "per_table_per_gpu_sent": int(n * 256 / world),
                                 "tables_per_step": 2 * k + 2,
                                 "note": "every produced row (256 B) is stored once by its owner into all copies (NVLS multicast: "
                                         "the switch replicates); tables per step: y_0..y_{K-1}, final^, z_0..z_{K-1}"}
    if world > 1 and shape == "ml25m":
        try:
            out["eval_full_rank_c4_sharded"] = block_eval_c4_sharded(nu, ni, ops.edge_index, world, rank, args.steps,
                                                                     max_over_ranks, trainer.comm.barrier)
        except Exception as exc:                                # every rank takes the same path: no rank is left waiting
            out["eval_full_rank_c4_sharded"] = {"error": repr(exc)[:300]}
    if args.check:
        out["parity"] = parity_check(trainer, ops, nu, ni, train, k, dev, world, rank)
    return out, dict(nu=nu, ni=ni, train=train, k=k, dev=dev, world=world, rank=rank, trainer=trainer, ops=ops, shape=shape,
                     graph=graph)


def parity_check(trainer, ops, nu, ni, train, k, dev, world, rank):
    """One step from freshly initialised weights on identical negatives against the float64 restatement (rank 0's GPU):
    loss, dL/dE0 (assembled from every rank's owned rows) and the propagated final rows, normwise."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import fp64_ref
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    trainer.user_w.copy_(u0.to(dev)); trainer.item_w.copy_(i0.to(dev))
    ops.m.zero_(); ops.v.zero_(); ops.step_count.zero_()
    trainer.drop_graph()
    fin = trainer.propagate_only().clone()
    gen = torch.Generator().manual_seed(29)
    neg = torch.randint(0, ni, (ops.num_triplets,), generator=gen).to(dev)
    loss = float(trainer.step(neg))
    grad = ops.grad.clone()
    if world > 1:                                   # every rank contributes its owned rows (the others are stale)
        own = torch.zeros(nu + ni, dtype=torch.bool, device=dev)
        for rb, re in trainer.segs:
            own[rb:re] = True
        grad[~own] = 0
        dist.all_reduce(grad)
    res = None
    if rank == 0:
        e0 = torch.cat([u0, i0]).to(dev).double()
        rl, rgrad, rfinal = fp64_ref.step_loss_and_grad(train.to(dev), e0, k, nu, neg)

        def nw(a, b):
            return float((a.double() - b).abs().max() / b.abs().max())
        res = {"against": "float64 restatement of the reference's op sequence on the device (tests/fp64_ref.py)",
               "loss": loss, "loss_fp64": rl, "loss_rel": abs(loss - rl) / abs(rl),
               "final_normwise": nw(fin, rfinal), "grad_e0_normwise": nw(grad, rgrad), "tolerance": 1e-5}
        res["ok"] = bool(res["loss_rel"] < 1e-5 and res["final_normwise"] < 1e-5 and res["grad_e0_normwise"] < 1e-5)
    if world > 1:
        dist.barrier()
    return res


# --------------------------------------------------------------------------------------------
# named blocks of the N = 1 line
# --------------------------------------------------------------------------------------------

def block_cluster_gcn_c2(nu, ni, train, k, dev, steps, warmup, with_cpu):
    """BASELINE configs[1]: one Cluster-GCN epoch (100 METIS parts of the ML-25M-shaped train graph) through train():
    device-resident batches (`value`), host batches (`e2e`), the persistent step kernel's roofline, the CPU oracle."""
    from lgcn_b200.data.dataset_handler import ClusterData, ClusterLoader, Data
    from lgcn_b200.models.light_gcn import LightGCN
    from lgcn_b200.utils import train_test as tt
    n = nu + ni
    cluster = load_partition(train, n, "ml25m")
    cd = ClusterData(Data(edge_index=train.to(dev), num_nodes=n), NUM_PARTS, cluster=cluster)
    parts = list(cd.parts)
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    model = LightGCN(nu, ni, num_layers=k).to(dev)
    with torch.no_grad():
        model.user_embedding.weight.copy_(u0)
        model.item_embedding.weight.copy_(i0)
    opt = tt.FusedAdam(model)
    graphs = [model.graph(d.edge_index) for d in parts]
    live = [d for d, gr in zip(parts, graphs) if gr.num_triplets > 0]
    grs = [gr for gr in graphs if gr.num_triplets > 0]
    edges_per_epoch = int(sum(d.edge_index.shape[1] for d in live))
    loader = ClusterLoader(live, shuffle=True)
    torch.manual_seed(0)
    for _ in range(warmup):
        tt.train(model, opt, loader, dev)
    torch.cuda.synchronize()
    a, z = event_pair()
    a.record()
    for _ in range(steps):
        last = tt.train(model, opt, loader, dev)
    z.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(z) / steps
    host_parts = [Data(edge_index=d.edge_index.cpu().pin_memory(), num_nodes=n) for d in live]

    class HostLoader:
        """What the reference's DataLoader hands to train(): batches that live on the HOST (pinned);
        train() uploads them (utils/train_test.py:87 `batch.to(device)`) every epoch."""
        def __iter__(self):
            return iter(host_parts)

    model._graphs.capacity = 4
    for _ in range(2):
        tt.train(model, opt, HostLoader(), dev)
    torch.cuda.synchronize()
    a, z = event_pair()
    a.record()
    for _ in range(steps):
        tt.train(model, opt, HostLoader(), dev)
    z.record()
    torch.cuda.synchronize()
    e2e_ms = a.elapsed_time(z) / steps
    peak, _ = measured_peak_gbs()
    step_bytes = 0
    for gr in grs:
        touched = gr.num_active + min(gr.num_triplets, ni)
        step_bytes += touched * 256 * 6 + gr.num_active * 256 * (2 * (2 * k + 3)) + gr.num_edges * 4 * (2 * k + 4)
    roof = {"kernel": "epoch_kernel (lgcn_train_steps_sparse: all sparse steps of an epoch in ONE persistent launch)",
            "bound": "hbm", "achieved": step_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "algorithmic_bytes_per_launch": step_bytes, "avg_launch_ms": ms, "traffic": ncu_traffic("epoch_kernel"),
            "note": "latency-bound by construction: ~2 k rows / ~7 k edges per step and 2K+2 device-wide barriers per "
                    "step; duration = the whole epoch (kernel + end-of-epoch flush)"}
    roof["frac"] = roof["achieved"] / peak
    blk = {"workload": f"C2 Cluster-GCN training epoch: {NUM_PARTS} METIS parts ({len(live)} non-empty, {edges_per_epoch} "
                       f"intra-cluster edges), K={k}, fwd+BPR+bwd+clip+Adam per batch",
           "ms_per_epoch": ms, "value": edges_per_epoch / (ms * 1e-3), "unit": UNIT,
           "e2e": {"value": edges_per_epoch / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_epoch": e2e_ms,
                   "h2d_bytes_per_step": int(sum(hp.edge_index.numel() * 8 for hp in host_parts)),
                   "d2h_bytes_per_step": 4 * len(host_parts),
                   "note": "train() over HOST (pinned) batches: every epoch uploads all edge lists, rebuilds every "
                           "batch's normalisation / CSR (one batched K0b call) and reads the losses back"},
           "roofline": roof, "final_epoch_loss": last, "gpu_launches_per_epoch": 2}
    try:                                                            # f4: the GPU partitioner next to METIS (set-up stage, untimed above)
        from lgcn_b200.data import partition_gpu as pg
        tr_dev = train.to(dev)
        pg.partition(tr_dev, n, NUM_PARTS, nu)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, st = pg.partition(tr_dev, n, NUM_PARTS, nu)
        torch.cuda.synchronize()
        blk["gpu_partitioner"] = {"ms": (time.perf_counter() - t0) * 1e3, "intra_edge_share": st["intra_edge_share"],
                                  "metis_intra_edge_share": edges_per_epoch / train.shape[1], "max_part": st["max_part"],
                                  "starts": st["starts"],
                                  "note": "balanced label propagation on the device (data/partition_gpu.py), wall clock incl. its "
                                          "host syncs; METIS on the host takes ~75 s for this graph; the epoch above trains on the "
                                          "METIS parts (the reference's partitions)"}
        del tr_dev
    except Exception as exc:
        blk["gpu_partitioner"] = {"error": repr(exc)[:300]}
    if with_cpu:
        batches = [b for b in cluster_batches_cpu(train, cluster, n) if int((b[0] < nu).sum()) > 0]
        blk["cpu_baseline"] = cpu_cluster_epoch(nu, ni, batches, k)
    del model, opt, cd
    return blk


def block_spmm(nu, ni, train, k, dev):
    """The SpMM figures of the metric: K-layer fused propagation over the whole train graph (single C-ABI call)."""
    from lgcn_b200 import _lib
    from lgcn_b200.models.light_gcn import LightGCN
    L = _lib.lib()
    peak, _ = measured_peak_gbs()
    model = LightGCN(nu, ni, num_layers=k).to(dev)
    g = model.graph(train.to(dev))
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
        a, z = event_pair()
        a.record(); run(); z.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(z))
    ms = float(np.median(ts))
    b_layer = 2 * n * 256 + 8 * e + 4 * (n + 1)
    gather = e * 264 + n * 256 + 4 * (n + 1)
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
        a, z = event_pair()
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


def block_eval_c4(nu, ni, train_dev, test_dev, steps):
    """BASELINE configs[3]: full-rank evaluation, all users x all items, train-edge mask, top-20."""
    from lgcn_b200.utils import recommend as rec
    dev = train_dev.device
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    ue, ie = u0.to(dev), i0.to(dev)
    ptr, idx = rec.exclusion_csr(train_dev, nu)
    k = 20

    def timed(algo):
        def run():
            return rec.score_topk(ue, ie, k, True, ptr, idx, algo=algo)
        for _ in range(2):
            run()
        ts = []
        for _ in range(max(3, min(steps, 10))):
            a, z = event_pair()
            a.record(); run(); z.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(z))
        return float(np.median(ts))
    ms = timed(rec.SCORE_TENSOR)
    ms_ffma = timed(rec.SCORE_FFMA)
    m = rec.full_rank_eval(ue, ie, train_dev, test_dev, nu, k=k)
    flop = 2.0 * nu * ni * 64
    tf32_peak = measured_bf16_tflops() / 2
    return {"workload": f"C4 full-rank eval {nu} x {ni} x 64, train-edge mask, top-{k}",
            "ms": ms, "scores_per_s": nu * ni / (ms * 1e-3),
            "useful_tflops": flop / (ms * 1e-3) / 1e12, "issued_tf32_tflops": 3 * flop / (ms * 1e-3) / 1e12,
            "roofline": {"kernel": "score_topk_tc_kernel (tcgen05.mma kind::tf32, 3-term hi/lo split)", "bound": "tensor",
                         "achieved": 3 * flop / (ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": 3 * flop / (ms * 1e-3) / 1e12 / tf32_peak,
                         "peak_source": "half the measured dense bf16 peak (MEASURED_PEAKS.json)", "traffic": None,
                         "tensor_pipe_cycles_active_pct_ncu": (json.load(open(os.path.join(REPO, "profiles", "roofline_traffic.json")))
                                                               .get("score_topk_tc_one_wave_r2", {}).get("tensor_pipe_cycles_active_pct")
                                                               if os.path.exists(os.path.join(REPO, "profiles", "roofline_traffic.json")) else None)},
            "math": "tcgen05.mma kind::tf32, 3-term hi/lo split (fp32-level accuracy), fp32 accumulate in TMEM",
            "ffma_kernel_ms": ms_ffma, "recall@20": m["recall"], "ndcg@20": m["ndcg"],
            "users_with_test_items": m["users"], "note": "random-init embeddings: recall/NDCG are chance level"}


def block_eval_c4_sharded(nu, ni, train_dev, world, rank, steps, max_over_ranks, barrier):
    """BASELINE configs[3] over the N GPUs of the job: rank r scores its user range against the
    replicated item table (train-edge mask, top-20); no collective on the data path.  Time = max over ranks."""
    from lgcn_b200.utils import recommend as rec
    dev = train_dev.device
    u0, i0 = synthetic.init_embeddings(nu, ni, 64, 0)
    ue, ie = u0.to(dev), i0.to(dev)
    ptr, idx = rec.exclusion_csr(train_dev, nu)
    lo, hi = rec.user_ranges(nu, world)[rank]

    def run():
        return rec.score_topk(ue, ie, 20, True, ptr, idx, lo, hi)
    for _ in range(2):
        run()
    torch.cuda.synchronize(); barrier()
    reps = max(3, min(steps, 10))
    a, z = event_pair()
    a.record()
    for _ in range(reps):
        top, _ = run()
    z.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(z) / reps)
    digest = max_over_ranks(float(top.to(torch.float64).sum())) if world == 1 else None
    flop = 2.0 * nu * ni * 64
    return {"workload": f"C4 full-rank eval {nu} x {ni} x 64, train-edge mask, top-20, users split over {world} GPU(s) "
                        f"in equal ranges (multiples of the 128-user CTA tile)",
            "ms": ms, "scores_per_s": nu * ni / (ms * 1e-3), "useful_tflops": flop / (ms * 1e-3) / 1e12,
            "issued_tf32_tflops": 3 * flop / (ms * 1e-3) / 1e12, "users_rank0": hi - lo, "digest": digest}


# --------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------

def run_reference(args):
    """The reference's CPU implementation of the path = the oracle port (PyG / torch_sparse are not installable
    offline and /root/reference does not exist on the GPU box), all host threads, on OUR arm's workload: each step is
    a bounded sample of the full-graph training step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    shape = SHAPE_OF[args.workload]
    g = synthetic.make_graph(shape, seed=0)
    k = synthetic.SHAPES[shape][3]
    train = g.edges("train")
    cpu, sec = cpu_full_graph_sample(g.num_users, g.num_items, train, k, args.steps, args.warmup, budget_s=150.0)
    return {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload.upper()} full-graph training step (= full-graph epoch), {shape} shape: "
                                   f"U={g.num_users} I={g.num_items} E_train={train.shape[1]} directed, K={k}, dim=64, "
                                   f"fwd+BPR+bwd+clip+Adam, node-range sharded over N GPU(s)",
                       "reference": "CPU path = oracle port of the reference's PyG gather / scatter_add op sequence + BPR + "
                                    "clip + Adam (torch CPU, all host threads); one step = the bounded sample named in "
                                    "cpu_baseline.sample; ms_per_step is the measured time of that sample step"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c5", "c2", "c4"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the C2 / C4 / SpMM blocks")
    ap.add_argument("--check", action="store_true", help="after the timing: one more step against the float64 restatement "
                    "(tests/fp64_ref.py) on rank 0's GPU -> `parity` block")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload in ("c2", "c4"):
            args.workload = "c3"
        out = run_reference(args)
        if out is not None:
            print(json.dumps(out))
        return
    if args.workload in ("c2", "c4"):                          # standalone blocks (development aid)
        dev = torch.device("cuda:0")
        torch.cuda.set_device(dev)
        g = synthetic.make_graph("ml25m", seed=0)
        if args.workload == "c2":
            out = block_cluster_gcn_c2(g.num_users, g.num_items, g.edges("train"), 3, dev, args.steps, args.warmup,
                                       not args.no_cpu)
        else:
            out = block_eval_c4(g.num_users, g.num_items, g.edges("train").to(dev), g.edges("test").to(dev), args.steps)
        print(json.dumps(out))
        return
    out, ctx = run_full_graph(args)
    import torch.distributed as dist
    if ctx["world"] == 1 and ctx["shape"] == "ml25m":
        nu, ni, train, k, dev = ctx["nu"], ctx["ni"], ctx["train"], ctx["k"], ctx["dev"]
        del ctx["trainer"], ctx["ops"]
        torch.cuda.empty_cache()
        if not args.no_cpu:
            out["cpu_baseline"], _ = cpu_full_graph_sample(nu, ni, train, k, 2, 1, budget_s=25.0)
        if not args.no_extras:
            for name, fn in (("spmm_full_graph", lambda: block_spmm(nu, ni, train, k, dev)),
                             ("cluster_gcn_epoch_c2", lambda: block_cluster_gcn_c2(nu, ni, train, k, dev, args.steps,
                                                                                   args.warmup, not args.no_cpu)),
                             ("eval_full_rank_c4", lambda: block_eval_c4(
                                 nu, ni, train.to(dev), ctx["graph"].edges("test").to(dev), args.steps))):
                try:
                    out[name] = fn()
                except Exception as exc:                        # an extra is never the reason for a missing line
                    out[name] = {"error": repr(exc)[:300]}
                torch.cuda.empty_cache()
    if ctx["world"] > 1:
        dist.barrier()
        dist.destroy_process_group()
    if ctx["rank"] == 0:
        emit_line(out)


if __name__ == "__main__":
    main()
