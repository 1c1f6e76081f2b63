"""Node-range sharded full-graph LightGCN training step (BASELINE configs C3 / C5).

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Every rank holds the full
CSR pair of the edge list and executes only the warp tasks of the rows it OWNS: a contiguous user
range plus a contiguous item range, each balanced by edge count, so that every rank gets its share
of SpMM rows of both kinds and of BPR triplets (triplets follow their user).  Rows are
owner-computed (no float atomics across ranks); the exchange steps are

    forward   all-gather of the owned slabs of y_k = dis (.) x_k before layer k+1     (K all-gathers)
              all-gather of the final embeddings + their inverse norms (BPR reads remote rows)
    BPR       all-reduce(sum) of dL/dfinal [N,64] and of the negative-sample histogram
    backward  all-gather of z_j = dis (.) h_j before layer j+1                       (K-1 all-gathers)
    clip      all-reduce of (softplus sum, regulariser sum, ||grad||^2)              (3 doubles)

Weights and Adam state are updated for owned rows only -- no parameter all-reduce; ``gather_weights``
assembles the full tables (checkpointing).  With world_size 1 the result equals ``lgcn_train_step``.

The orchestration is backend-agnostic (``ops``): ``CudaOps`` drives the C ABI; the CPU test-suite
injects a torch stand-in to check the sharded algorithm over gloo with world_size 2.
"""
from __future__ import annotations

import os
from ctypes import byref
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

DIM = 64


# ----------------------------------------------------------------------------------------------
# ownership plan
# ----------------------------------------------------------------------------------------------

def balanced_boundaries(weight: torch.Tensor, parts: int) -> List[int]:
    """Split [0,len) into ``parts`` contiguous ranges of near-equal total weight (prefix-sum cut)."""
    n = weight.numel()
    if n == 0:
        return [0] * (parts + 1)
    pre = torch.cumsum(weight.to(torch.float64), 0)
    total = float(pre[-1])
    cuts = [0]
    for p in range(1, parts):
        cuts.append(int(torch.searchsorted(pre, torch.tensor(total * p / parts, dtype=torch.float64))))
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


@dataclass
class ShardPlan:
    """Row ownership: rank r owns user rows [user_ptr[r], user_ptr[r+1]) and item rows
    [item_ptr[r], item_ptr[r+1]) (global node ids; items start at num_users)."""
    num_users: int
    num_items: int
    user_ptr: List[int]
    item_ptr: List[int]

    @property
    def world(self) -> int:
        return len(self.user_ptr) - 1

    def segments(self, rank: int) -> List[Tuple[int, int]]:
        return [(self.user_ptr[rank], self.user_ptr[rank + 1]), (self.item_ptr[rank], self.item_ptr[rank + 1])]

    @staticmethod
    def build(in_deg: torch.Tensor, out_deg: torch.Tensor, num_users: int, world: int, row_cost: int = 8) -> "ShardPlan":
        w = (in_deg + out_deg + row_cost).cpu()
        n = w.numel()
        up = balanced_boundaries(w[:num_users], world)
        ip = [num_users + c for c in balanced_boundaries(w[num_users:], world)]
        return ShardPlan(num_users, n - num_users, up, ip)


# ----------------------------------------------------------------------------------------------
# collectives (plumbing)
# ----------------------------------------------------------------------------------------------

class Comm:
    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        self.backend = dist.get_backend(group) if self.on else "none"

    def allgather_rows(self, buf: torch.Tensor, ptr: Sequence[int]) -> None:
        """In place: every rank contributes buf[ptr[r]:ptr[r+1]] and receives the other slabs."""
        if self.world == 1:
            return
        views = [buf[ptr[r]:ptr[r + 1]] for r in range(self.world)]
        sizes = {v.shape[0] for v in views}
        if self.backend == "nccl":
            if len(sizes) == 1:
                lo, hi = ptr[0], ptr[-1]
                dist.all_gather_into_tensor(buf[lo:hi], views[self.rank], group=self.group)
            else:
                dist.all_gather(views, views[self.rank], group=self.group)   # uneven: grouped broadcasts
        else:                                                                # gloo: equal sizes only
            for r in range(self.world):
                if views[r].numel():
                    dist.broadcast(views[r], src=dist.get_global_rank(self.group, r) if self.group else r,
                                   group=self.group)

    def allreduce(self, t: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def barrier(self) -> None:
        if self.world > 1:
            dist.barrier(group=self.group)


# ----------------------------------------------------------------------------------------------
# CUDA backend
# ----------------------------------------------------------------------------------------------

class CudaOps:
    """The product backend: each method is one C-ABI call on the rank's row / task ranges.

    p2p=True (default when world > 1): the exchanged tables (y_k, z_j, final, rnorm) live in one
    symmetric-memory region and the kernels store every produced row straight into all ranks' copies
    (NVLS multicast store, or per-peer stores over NVLink) -- the all-gather is fused into the SpMM
    epilogue and only a cross-rank barrier separates layers.  p2p=False: NCCL all-gathers between
    kernels (also the fallback when symmetric memory cannot be set up)."""

    def __init__(self, edge_index: torch.Tensor, num_users: int, num_items: int, num_layers: int,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0,
                 p2p: Optional[bool] = None, group=None):
        from . import _lib
        self._lib = _lib
        self.L = _lib.lib()
        self.dev = edge_index.device
        self.g = _lib.Graph(edge_index, num_users, num_items)
        self.nu, self.ni, self.n, self.k = num_users, num_items, num_users + num_items, num_layers
        f32 = dict(dtype=torch.float32, device=self.dev)
        n = self.n
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if p2p is None:
            p2p = world > 1 and os.environ.get("LGCN_P2P", "1") != "0"
        nz = min(2, max(num_layers - 1, 0))
        total = (num_layers + nz + 1) * n * DIM + n + 64          # + barrier flags (int32 view of the tail)
        self.p2p, self.peers, self.hdl, self.p2p_error = False, None, None, None
        flat = None
        if p2p and world > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                flat = symm_mem.empty(total, dtype=torch.float32, device=self.dev)
                self.hdl = symm_mem.rendezvous(flat, group if group is not None else dist.group.WORLD)
                c = _lib.CPeers()
                c.world, c.rank = self.hdl.world_size, self.hdl.rank
                mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
                if os.environ.get("LGCN_P2P_MULTICAST", "1") == "0":
                    mc = 0
                c.mc_base = mc if mc else None
                for i, ptr in enumerate(self.hdl.buffer_ptrs):
                    c.base[i] = ptr
                self.peers, self.p2p, self.multicast = c, True, bool(mc)
            except Exception as exc:           # no fabric / VMM support: NCCL all-gathers instead
                self.p2p_error, flat = f"{type(exc).__name__}: {exc}", None
        if flat is None:
            flat = torch.empty(total, **f32)
        flat.zero_()
        off = 0

        def take(rows_by_cols):
            nonlocal off
            t = flat[off:off + rows_by_cols]
            off += rows_by_cols
            return t
        self.y = [take(n * DIM).view(n, DIM) for _ in range(num_layers)]          # y_0 .. y_{K-1}
        self.z = [take(n * DIM).view(n, DIM) for _ in range(nz)]
        self.final = take(n * DIM).view(n, DIM)
        self.rnorm = take(n)
        self.flags = take(64).view(torch.int32)                    # [0:8] = per-rank arrival epochs
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._flat = flat
        if self.p2p:
            torch.cuda.synchronize(self.dev)
            self.hdl.barrier(channel=0)            # everyone has zeroed its region before the first flag lands
            torch.cuda.synchronize(self.dev)
        self.G = torch.zeros(n, DIM, **f32)
        self.grad = torch.zeros(n, DIM, **f32)
        self.neg_count = torch.zeros(num_items, dtype=torch.int32, device=self.dev)
        self.scratch = torch.empty(2 * max(self.g.num_triplets, 1), **f32)
        self.accum = torch.zeros(4, dtype=torch.float64, device=self.dev)
        self.loss = torch.zeros(1, **f32)
        self.m, self.v = torch.zeros(n, DIM, **f32), torch.zeros(n, DIM, **f32)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=self.dev)
        c = _lib.CAdam()
        c.lr, c.beta1, c.beta2, c.eps, c.max_norm = lr, betas[0], betas[1], eps, max_norm
        c.step, c.m, c.v = self.step_count.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
        self.adam = c
        self.in_rows = self.g.in_tasks.view(-1, 8)[: self.g.c.n_in_tasks, 0].cpu()
        self.out_rows = self.g.out_tasks.view(-1, 8)[: self.g.c.n_out_tasks, 0].cpu()
        self.local = None            # set by bind_segments(): rank-local task lists (one launch per layer)

    def bind_segments(self, segs) -> None:
        """Task order is free, so the tasks of ALL row segments this rank owns are packed into one
        contiguous device list per direction: a layer is then ONE launch per rank (not one per
        segment), and the ranges are resolved once instead of per call."""
        import copy
        lib_ = self._lib
        parts_in, parts_out, n_user_out = [], [], 0
        for rb, re in segs:
            tb, te = self._tasks(self.in_rows, rb, re)
            parts_in.append(self.g.in_tasks.view(-1, 8)[tb:te])
            tb, te = self._tasks(self.out_rows, rb, re)
            parts_out.append(self.g.out_tasks.view(-1, 8)[tb:te])
            if re <= self.nu:
                n_user_out += te - tb
        self.loc_in = torch.cat(parts_in).contiguous().view(-1)
        self.loc_out = torch.cat(parts_out).contiguous().view(-1)
        c = lib_.CGraph.from_buffer_copy(self.g.c)
        c.in_tasks, c.out_tasks = self.loc_in.data_ptr(), self.loc_out.data_ptr()
        c.n_in_tasks, c.n_out_tasks = self.loc_in.numel() // 8, self.loc_out.numel() // 8
        c.n_out_user_tasks = n_user_out
        self.sched2 = torch.zeros(2, dtype=torch.int32, device=self.dev)
        c.sched = self.sched2.data_ptr()
        self.local = c
        self.segs = list(segs)
        self.all_active = self.g.num_active == self.n

    # -- graph facts the planner needs
    @property
    def num_triplets(self) -> int:
        return self.g.num_triplets

    def degrees(self):
        return self.g.in_degree(), self.g.out_degree()

    def _tasks(self, rows: torch.Tensor, rb: int, re: int) -> Tuple[int, int]:
        b = int(torch.searchsorted(rows, torch.tensor(rb, dtype=rows.dtype)))
        e = int(torch.searchsorted(rows, torch.tensor(re, dtype=rows.dtype)))
        return b, e

    def set_weights(self, user_w: torch.Tensor, item_w: torch.Tensor) -> None:
        self.uw, self.iw = user_w, item_w

    def _s(self):
        return self._lib.stream_ptr(self.dev)

    def _p(self):
        return byref(self.peers) if self.p2p else None

    def peer_barrier(self):
        """All ranks' rows of the table just produced have landed everywhere (stream-ordered)."""
        self._lib.check(self.L.lgcn_peer_barrier(byref(self.peers), self.flags.data_ptr(), self.epoch.data_ptr(),
                                                 self._s()))

    def step_begin(self):
        self._lib.check(self.L.lgcn_step_begin(byref(self.adam), self.accum.data_ptr(), self._s()))

    def prescale(self, rb, re):
        self._lib.check(self.L.lgcn_prescale(self.g.ref, self.uw.data_ptr(), self.iw.data_ptr(), rb, re,
                                             self.y[0].data_ptr(), self._p(), self._s()))

    def fwd_layer(self, k, rb, re):
        """Layer k over rows [rb,re); (rb,re) == (None,None): all of this rank's segments at once."""
        if rb is None:
            gref, tb, te, rb, re = byref(self.local), 0, self.local.n_in_tasks, 0, 0
        else:
            gref = self.g.ref
            tb, te = self._tasks(self.in_rows, rb, re)
        last = k == self.k
        ys = [self.y[i].data_ptr() if i < self.k else None for i in (1, 2, 3)]
        self._lib.check(self.L.lgcn_fwd_layer(
            gref, self.uw.data_ptr(), self.iw.data_ptr(), k, self.k, self.y[k - 1].data_ptr(),
            None if last else self.y[k].data_ptr(), ys[0], ys[1], ys[2],
            self.final.data_ptr() if last else None, self.rnorm.data_ptr() if last else None, tb, te, rb, re,
            self._p(), self._s()))

    def bpr(self, neg, urb, ure):
        if self.local is not None:       # user tasks are the prefix of the packed by-source list; the item pass
            gref, tb, te = byref(self.bpr_graph()), 0, self.local.n_out_user_tasks     # walks the GLOBAL by-target list
        else:
            gref = self.g.ref
            tb, te = self._tasks(self.out_rows, urb, ure)
        self._lib.check(self.L.lgcn_bpr_fwd_bwd_range(
            gref, self.final.data_ptr(), self.rnorm.data_ptr(), neg.data_ptr(), self.G.data_ptr(),
            self.neg_count.data_ptr(), self.scratch.data_ptr(), self.accum.data_ptr(), tb, te, urb, ure, self._s()))

    def bpr_graph(self):
        """Packed by-source tasks (pass A: own users) + the global by-target tasks (pass B: all items,
        filtered by user range)."""
        if getattr(self, "_bpr_c", None) is None:
            c = self._lib.CGraph.from_buffer_copy(self.g.c)
            c.out_tasks, c.n_out_tasks, c.n_out_user_tasks = self.local.out_tasks, self.local.n_out_tasks, self.local.n_out_user_tasks
            c.sched = self.local.sched
            self._bpr_c = c
        return self._bpr_c

    def bwd_layer(self, j, rb, re, bpr_coeff):
        if rb is None:
            gref, tb, te, rb, re = byref(self.local), 0, self.local.n_out_tasks, 0, 0
        else:
            gref = self.g.ref
            tb, te = self._tasks(self.out_rows, rb, re)
        last = j == self.k
        zin = None if j == 1 else self.z[j & 1].data_ptr()
        zout = None if last else self.z[(j - 1) & 1].data_ptr()
        reg = 2.0 * bpr_coeff / (64.0 * self.g.num_triplets)
        self._lib.check(self.L.lgcn_bwd_layer(
            gref, self.G.data_ptr(), j, self.k, zin, zout, self.uw.data_ptr(), self.iw.data_ptr(),
            self.neg_count.data_ptr(), reg, self.grad.data_ptr() if last else None, self.accum.data_ptr(),
            tb, te, rb, re, self._p(), self._s()))

    def zbuf(self, j):
        return self.z[(j - 1) & 1]

    def clip_adam(self, rb, re, bpr_coeff):
        self._lib.check(self.L.lgcn_clip_adam_rows(
            byref(self.adam), self.uw.data_ptr(), self.iw.data_ptr(), self.nu, self.ni, self.grad.data_ptr(),
            self.accum.data_ptr(), self.g.num_triplets, bpr_coeff, self.loss.data_ptr(), rb, re, self._s()))


# ----------------------------------------------------------------------------------------------
# orchestration
# ----------------------------------------------------------------------------------------------

class ShardedTrainer:
    """Full-graph training step of utils/train_test.py:88-96, node-range sharded."""

    def __init__(self, ops, user_w: torch.Tensor, item_w: torch.Tensor, comm: Optional[Comm] = None,
                 bpr_coeff: float = 5e-3, plan: Optional[ShardPlan] = None):
        self.ops = ops
        self.comm = comm if comm is not None else Comm()
        self.bpr_coeff = bpr_coeff
        self.user_w, self.item_w = user_w, item_w
        ops.set_weights(user_w, item_w)
        ind, outd = ops.degrees()
        self.plan = plan if plan is not None else ShardPlan.build(ind, outd, ops.nu, self.comm.world)
        self.segs = self.plan.segments(self.comm.rank)
        self.k = ops.k
        # one launch per layer when the backend can pack this rank's segments and no row is edge-less
        self.packed = False
        if hasattr(ops, "bind_segments"):
            ops.bind_segments(self.segs)
            self.packed = ops.all_active
        self.layer_segs = [(None, None)] if self.packed else self.segs

    def _gather(self, buf: torch.Tensor, produced_by_kernel: bool = True) -> None:
        """Make every rank's copy of ``buf`` complete.  Tables produced by a p2p-enabled kernel are
        already being written into all copies; only a barrier is needed."""
        if produced_by_kernel and getattr(self.ops, "p2p", False):
            self.ops.peer_barrier()
            return
        self.comm.allgather_rows(buf, self.plan.user_ptr)
        self.comm.allgather_rows(buf, self.plan.item_ptr)

    def step(self, neg: torch.Tensor) -> torch.Tensor:
        """One step; ``neg`` is the FULL [P] negative vector (identical on every rank; each rank reads
        the entries of its own triplets).  Returns the loss as a device tensor (same on every rank)."""
        o, k = self.ops, self.k
        o.step_begin()
        for rb, re in self.segs:
            o.prescale(rb, re)
        self._gather(o.y[0])
        for layer in range(1, k + 1):
            for rb, re in self.layer_segs:
                o.fwd_layer(layer, rb, re)
            if layer < k:
                self._gather(o.y[layer])
        self._gather(o.final)
        if not getattr(o, "p2p", False):
            self._gather(o.rnorm)             # (p2p: same kernel, same barrier as `final`)
        urb, ure = self.segs[0]
        o.bpr(neg, urb, ure)
        self.comm.allreduce(o.G)
        self.comm.allreduce(o.neg_count)
        for j in range(1, k + 1):
            for rb, re in self.layer_segs:
                o.bwd_layer(j, rb, re, self.bpr_coeff)
            if j < k:
                self._gather(o.zbuf(j))
        self.comm.allreduce(o.accum)
        for rb, re in self.segs:
            o.clip_adam(rb, re, self.bpr_coeff)
        return o.loss

    def step_sampled(self, num_items: Optional[int] = None, use_graph: bool = True) -> torch.Tensor:
        """One step with the reference's negative sampling (uniform ``randint`` per triplet,
        utils/helpers.py:79-80) done on the device.  All ranks must hold the same torch CUDA RNG state
        (same ``torch.manual_seed``) so that they draw identical negatives.  From the 4th call on the
        whole step -- sampling, kernels, barriers, NCCL all-reduces -- is replayed as ONE CUDA graph
        launch per rank, which removes the ~25 host calls per step that dominate at 8 GPUs."""
        ni = self.plan.num_items if num_items is None else num_items
        p, dev = self.ops.num_triplets, self.user_w.device
        self._calls = getattr(self, "_calls", 0) + 1
        if not use_graph or not self.user_w.is_cuda or getattr(self, "_graph_failed", False):
            return self.step(torch.randint(0, ni, (p,), device=dev))
        if getattr(self, "_graph", None) is None:
            if self._calls <= 3:                       # eager warm-up (NCCL channels, allocator)
                return self.step(torch.randint(0, ni, (p,), device=dev))
            try:
                torch.cuda.synchronize(dev)
                self.comm.barrier()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._graph_loss = self.step(torch.randint(0, ni, (p,), device=dev))
                self._graph = g
            except Exception as exc:                   # capture not supported for some op: stay eager
                self._graph_failed, self._graph_error = True, f"{type(exc).__name__}: {exc}"
                torch.cuda.synchronize(dev)
                return self.step(torch.randint(0, ni, (p,), device=dev))
        self._graph.replay()
        return self._graph_loss

    def gather_weights(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Full, up-to-date tables on every rank (for state_dict / best_model.pth)."""
        nu = self.plan.num_users
        full = torch.cat([self.user_w, self.item_w])
        self._gather(full, produced_by_kernel=False)
        return full[:nu].clone(), full[nu:].clone()

    def propagate_only(self) -> torch.Tensor:
        """Forward propagation alone (BASELINE config C3's per-layer timing): returns final [N,64]."""
        o, k = self.ops, self.k
        for rb, re in self.segs:
            o.prescale(rb, re)
        self._gather(o.y[0])
        for layer in range(1, k + 1):
            for rb, re in self.layer_segs:
                o.fwd_layer(layer, rb, re)
            if layer < k:
                self._gather(o.y[layer])
        self._gather(o.final)
        return o.final
