"""Node-range sharded full-graph LightGCN training step (BASELINE configs C3 / C5).

One process per GPU (``torch.distributed``; NVLink/NVSwitch peer memory for the data path).  Rank r OWNS a
contiguous user range plus a contiguous item range, each balanced by work, and holds only its SHARD of the edge
list: the edges with an owned endpoint (every edge has one user and one item endpoint, so it lives on at most two
ranks).  From the shard it builds its own CSR pair (K0); the rows it owns are complete in it, the others are never
executed.  Everything is owner-computed -- no float atomics, no reduction of a gradient table across ranks:

    forward   y_k = dis (.) x_k rows are stored into EVERY rank's copy from the SpMM epilogue (fused all-gather:
              one NVLS multicast store, or `world` peer stores), a peer barrier separates layers   (K+1 barriers)
    BPR       the user owner walks its users' triplets (loss, dL/dfinal[user]); the item owner walks the item's
              in-edges (positive role) and the bucket of this step's negatives that hit the item (negative role);
              per-triplet scalars are recomputed from the rows on both sides, so only embedding rows are ever read
              remotely.  Epilogues store z_0 = dis (.) dL/dfinal into every copy               (1 barrier)
    backward  K transpose-SpMM layers (Horner), z_j stored into every copy                     (K-1 barriers)
    clip      the three global sums ride on a barrier (lgcn_peer_allreduce4)                    (1 barrier)
    Adam      owned rows only -- no parameter all-reduce; ``gather_weights`` assembles full tables (checkpoints).

With world_size 1 the same code runs without exchange (and passes the triplet scalars instead of recomputing them).
When symmetric memory cannot be set up the tables are exchanged by NCCL all-gathers between kernels instead.

The orchestration is backend-agnostic (``ops``): ``CudaOps`` drives the C ABI; the CPU test-suite injects a torch
stand-in to check the sharded algorithm over gloo with world_size 2.
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
        """Cost of a row = the rows its owner gathers for it per step: 2K SpMM gathers per incident edge, BPR gathers
        per triplet role (a user plays one role per out-edge, an item one per in-edge plus, on average, P/I sampled
        negatives), plus a constant per row."""
        ind, outd = in_deg.cpu().double(), out_deg.cpu().double()
        n = ind.numel()
        num_items = n - num_users
        w = ind + outd + row_cost
        p = float(outd[:num_users].sum())
        wi = w[num_users:] + (p / num_items if num_items else 0.0)
        up = balanced_boundaries(w[:num_users], world)
        ip = [num_users + c for c in balanced_boundaries(wi, world)]
        return ShardPlan(num_users, num_items, up, ip)


@dataclass
class EdgeShard:
    """The part of an edge list a rank needs: edges with an owned endpoint, in their original relative order, plus the
    GLOBAL triplet number (rank among ALL user->movie edges, utils/helpers.py:98-99) of its user->movie edges."""
    edges: torch.Tensor            # [2, E_r] int64, global node ids
    trip_global: Optional[torch.Tensor]   # [P_r] int32, None = identity (whole list)
    num_triplets: int              # global P
    num_edges: int                 # global E

    @staticmethod
    def build(edge_index: torch.Tensor, plan: ShardPlan, rank: int) -> "EdgeShard":
        row, col = edge_index[0], edge_index[1]
        is_trip = row < plan.num_users
        p = int(is_trip.sum())
        if plan.world == 1:
            return EdgeShard(edge_index, None, p, edge_index.shape[1])
        (ub, ue), (ib, ie) = plan.segments(rank)
        own_r = ((row >= ub) & (row < ue)) | ((row >= ib) & (row < ie))
        own_c = ((col >= ub) & (col < ue)) | ((col >= ib) & (col < ie))
        keep = own_r | own_c
        gt = torch.cumsum(is_trip.to(torch.int32), 0, dtype=torch.int32) - 1
        return EdgeShard(edge_index[:, keep].contiguous(), gt[keep & is_trip].contiguous(), p, edge_index.shape[1])


# ----------------------------------------------------------------------------------------------
# collectives (plumbing; the CUDA data path exchanges through peer memory instead)
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
    """The product backend: each method is one C-ABI call over the rows this rank owns.

    ``edge_index``: the FULL [2,E] int64 edge list (CPU or CUDA tensor; the same on every rank) -- ``bind`` keeps only
    this rank's shard on the device.  p2p (default when world > 1): the exchanged tables live in one symmetric-memory
    region and the kernels store every produced row straight into all ranks' copies; p2p=False: NCCL all-gathers
    between kernels (also the fallback when symmetric memory cannot be set up)."""

    def __init__(self, edge_index: torch.Tensor, num_users: int, num_items: int, num_layers: int,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0,
                 p2p: Optional[bool] = None, group=None, device: Optional[torch.device] = None):
        from . import _lib
        self._lib = _lib
        self.L = _lib.lib()
        self.dev = torch.device(device) if device is not None else (
            edge_index.device if edge_index.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        self.edge_index = edge_index
        self.group = group
        self.nu, self.ni, self.n, self.k = num_users, num_items, num_users + num_items, num_layers
        self.hyper = (lr, betas, eps, max_norm)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.want_p2p = (self.world > 1 and os.environ.get("LGCN_P2P", "1") != "0") if p2p is None else (p2p and self.world > 1)
        self.p2p, self.peers, self.hdl, self.p2p_error, self.multicast = False, None, None, None, False
        self.g = None

    # -- graph facts the planner needs (global)
    def degrees(self):
        ei = self.edge_index
        return torch.bincount(ei[1], minlength=self.n), torch.bincount(ei[0], minlength=self.n)

    @property
    def num_triplets(self) -> int:
        return self.P

    # -- setup ---------------------------------------------------------------------------------------
    def bind(self, plan: ShardPlan, rank: int) -> None:
        """Shard the edge list, build this rank's CSR pair, allocate the tables, index the triplets."""
        self.plan, self.rank = plan, rank
        self.segs = plan.segments(rank)
        ei = self.edge_index
        shard = EdgeShard.build(ei, plan, rank)
        self.P, self.E = shard.num_triplets, shard.num_edges
        self.shard = shard
        self._alloc()
        self.load_shard(shard.edges.to(self.dev), None if shard.trip_global is None else shard.trip_global.to(self.dev))

    def _alloc(self) -> None:
        lib_, n, k, dev = self._lib, self.n, self.k, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        p = max(self.P, 1)
        trip_words = 2 * p + (2 * p) % 2
        # exchanged: y_0..y_{K-1}, z ping-pong (2), zg, final, rnorm, trip_user/trip_pos, accum slots, barrier flags
        total = (k + 4) * n * DIM + (n + n % 2) + trip_words + 64 + 64
        flat = None
        if self.want_p2p:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                flat = symm_mem.empty(total, dtype=torch.float32, device=dev)
                self.hdl = symm_mem.rendezvous(flat, self.group if self.group is not None else dist.group.WORLD)
                c = lib_.CPeers()
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

        def take(count):
            nonlocal off
            t = flat[off:off + count]
            off += count
            return t
        self.y = [take(n * DIM).view(n, DIM) for _ in range(k)]                   # y_0 .. y_{K-1}
        self.z = [take(n * DIM).view(n, DIM) for _ in range(2)]
        self.zg = take(n * DIM).view(n, DIM)                                      # z_0 = dis (.) dL/dfinal
        self.final = take(n * DIM).view(n, DIM)
        self.rnorm = take(n + n % 2)[:n]
        trip = take(trip_words).view(torch.int32)
        self.trip_user, self.trip_pos = trip[:p], trip[p:2 * p]
        self.slots = take(64).view(torch.float64)                                 # [8][4] doubles
        self.flags = take(64).view(torch.int32)                                   # [0:8] = per-rank arrival epochs
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self._flat = flat
        if self.p2p:
            torch.cuda.synchronize(dev)
            self.hdl.barrier(channel=0)            # everyone has zeroed its region before the first store lands
            torch.cuda.synchronize(dev)
        self.G = torch.zeros(n, DIM, **f32)        # dL/dfinal: owned rows only, never exchanged
        self.grad = torch.zeros(n, DIM, **f32)
        self.neg_count = torch.zeros(self.ni, dtype=torch.int32, device=dev)
        self.accum = torch.zeros(4, dtype=torch.float64, device=dev)
        self.loss = torch.zeros(1, **f32)
        self.m, self.v = torch.zeros(n, DIM, **f32), torch.zeros(n, DIM, **f32)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        lr, betas, eps, max_norm = self.hyper
        c = lib_.CAdam()
        c.lr, c.beta1, c.beta2, c.eps, c.max_norm = lr, betas[0], betas[1], eps, max_norm
        c.step, c.m, c.v = self.step_count.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
        self.adam = c
        # BPR workspace: buckets of the step's negatives over the owned items (+ scalars when nothing is sharded)
        (_, _), (ib, ie) = self.segs
        ni_own = ie - ib
        i32 = dict(dtype=torch.int32, device=dev)
        self.bucket_ptr = torch.zeros(ni_own + 2, **i32)
        self.bucket_cursor = torch.zeros(ni_own + 1, **i32)
        self.bucket = torch.empty(p, **i32)
        self.scalars = torch.empty(4 * p, **f32) if self.world == 1 else None
        self.bpr_sched = torch.zeros(2, **i32)
        w = lib_.CBprOwnerWs()
        w.trip_user, w.trip_pos = self.trip_user.data_ptr(), self.trip_pos.data_ptr()
        w.bucket_ptr, w.bucket_cursor = self.bucket_ptr.data_ptr(), self.bucket_cursor.data_ptr()
        w.bucket, w.bucket_cap = self.bucket.data_ptr(), p
        w.scalars = None if self.scalars is None else self.scalars.data_ptr()
        w.sched = self.bpr_sched.data_ptr()
        self.bpr_ws = w

    def load_shard(self, edges_dev: torch.Tensor, trip_global_dev: Optional[torch.Tensor]) -> None:
        """K0 on this rank's shard (already on the device) + global triplet numbers + the rank-local task lists + the
        triplet index (exchanged).  Called by ``bind`` and again by callers that re-upload the edge list."""
        lib_ = self._lib
        g = lib_.Graph(edges_dev, self.nu, self.ni)
        if trip_global_dev is not None:
            if trip_global_dev.numel() != g.num_triplets:
                raise lib_.LgcnError(f"shard has {g.num_triplets} user->movie edges, trip_global {trip_global_dev.numel()}")
            lib_.check(self.L.lgcn_graph_remap_triplets(g.ref, trip_global_dev.data_ptr(), self._s()))
        self.g = g
        self.graph_version = getattr(self, "graph_version", 0) + 1   # a captured step holds the OLD arrays' addresses
        self._pack_tasks()
        self.index_triplets()

    def _pack_tasks(self) -> None:
        """Task order is free (a split row's partials are reduced in slot order by whichever task arrives last), so
        the tasks of both row segments this rank owns are packed into contiguous device lists, one launch per layer
        and rank.  Every kernel pulls tasks from its list in order with a device counter: the lists are sorted by
        DESCENDING edge count, so the last tasks a launch hands out are the cheapest ones and no warp is left alone
        with a 512-edge task while the rest of the chip idles.  BPR walks user rows (out list) and item rows (in
        list) separately: those two lists are kept apart, sorted the same way."""
        g, lib_ = self.g, self._lib
        in_t, out_t = g.in_tasks.view(-1, 8)[: g.c.n_in_tasks], g.out_tasks.view(-1, 8)[: g.c.n_out_tasks]
        bounds = torch.tensor([b for seg in self.segs for b in seg], dtype=torch.int32, device=self.dev)
        bi = torch.searchsorted(in_t[:, 0].contiguous(), bounds).tolist()
        bo = torch.searchsorted(out_t[:, 0].contiguous(), bounds).tolist()

        def by_size(t):
            order = torch.sort(t[:, 2] - t[:, 1], descending=True, stable=True)[1]
            return t[order]
        in_u, in_i = in_t[bi[0]:bi[1]], in_t[bi[2]:bi[3]]
        out_u, out_i = out_t[bo[0]:bo[1]], out_t[bo[2]:bo[3]]
        self.loc_in = by_size(torch.cat([in_u, in_i])).contiguous().view(-1)          # SpMM forward: all owned rows
        self.loc_out = by_size(torch.cat([out_u, out_i])).contiguous().view(-1)       # SpMM backward
        self.bpr_users = by_size(out_u).contiguous().view(-1)                         # BPR user pass
        self.bpr_items = by_size(in_i).contiguous().view(-1)                          # BPR item pass
        self.sched2 = torch.zeros(2, dtype=torch.int32, device=self.dev)

        def view(in_list, out_list):
            c = lib_.CGraph.from_buffer_copy(g.c)
            c.in_tasks, c.out_tasks = in_list.data_ptr(), out_list.data_ptr()
            c.n_in_tasks, c.n_out_tasks = in_list.numel() // 8, out_list.numel() // 8
            c.n_in_user_tasks, c.n_out_user_tasks = 0, c.n_out_tasks
            c.sched = self.sched2.data_ptr()
            return c
        self.local = view(self.loc_in, self.loc_out)
        self.local_bpr = view(self.bpr_items, self.bpr_users)
        # owned rows without any incident edge need the dense "inactive row" kernels
        self.inactive_segs = [(rb, re) for rb, re in self.segs if re > rb and not bool(g.active[rb:re].all())]

    def index_triplets(self) -> None:
        if self.P == 0:
            return
        self._lib.check(self.L.lgcn_triplet_index(byref(self.local_bpr), 0, self.local_bpr.n_out_tasks,
                                                  self.trip_user.data_ptr(), self.trip_pos.data_ptr(), self._p(), self._s()))

    def set_weights(self, user_w: torch.Tensor, item_w: torch.Tensor) -> None:
        self.uw, self.iw = user_w, item_w

    def _s(self):
        return self._lib.stream_ptr(self.dev)

    def _p(self):
        return byref(self.peers) if self.p2p else None

    # -- exchange -----------------------------------------------------------------------------------
    def peer_barrier(self):
        """All ranks' rows of the table just produced have landed everywhere (stream-ordered)."""
        self._lib.check(self.L.lgcn_peer_barrier(byref(self.peers), self.flags.data_ptr(), self.epoch.data_ptr(),
                                                 self._s()))

    def allreduce_accum(self):
        self._lib.check(self.L.lgcn_peer_allreduce4(byref(self.peers), self.flags.data_ptr(), self.epoch.data_ptr(),
                                                    self.slots.data_ptr(), self.accum.data_ptr(), self._s()))

    # -- the step's stages ----------------------------------------------------------------------------
    def step_begin(self):
        self._lib.check(self.L.lgcn_step_begin(byref(self.adam), self.accum.data_ptr(), self._s()))

    def prescale(self):
        for rb, re in self.segs:
            self._lib.check(self.L.lgcn_prescale(self.g.ref, self.uw.data_ptr(), self.iw.data_ptr(), rb, re,
                                                 self.y[0].data_ptr(), self._p(), self._s()))

    def _fwd_call(self, gref, k, tb, te, rb, re, normalized):
        last = k == self.k
        ys = [self.y[i].data_ptr() if i < self.k else None for i in (1, 2, 3)]
        self._lib.check(self.L.lgcn_fwd_layer_ex(
            gref, self.uw.data_ptr(), self.iw.data_ptr(), k, self.k, self.y[k - 1].data_ptr(),
            None if last else self.y[k].data_ptr(), ys[0], ys[1], ys[2],
            self.final.data_ptr() if last else None, self.rnorm.data_ptr() if last else None, tb, te, rb, re,
            1 if (normalized and last) else 0, self._p(), self._s()))

    def fwd_layer(self, k, normalized=True):
        """normalized (training step): the last layer stores final / ||final|| into every copy of ``final`` and
        1/||final|| of the owned rows into ``rnorm``; False (inference): the plain final embeddings."""
        if k == self.k:
            for rb, re in self.inactive_segs:                    # only the inactive-row kernel runs for an empty task range
                self._fwd_call(self.g.ref, k, 0, 0, rb, re, normalized)
        self._fwd_call(byref(self.local), k, 0, self.local.n_in_tasks, 0, 0, normalized)

    def bpr_buckets(self, neg):
        """This step's negatives grouped by owned item.  Needs only ``neg``: the trainer runs it between the launch that
        produces y_0 and the barrier that waits for the peers' rows, so it hides behind that exchange."""
        (_, _), (ib, ie) = self.segs
        self._lib.check(self.L.lgcn_bpr_buckets(neg.data_ptr(), self.P, ib - self.nu, ie - self.nu,
                                                self.neg_count.data_ptr(), byref(self.bpr_ws), self._s()))
        self._buckets_for = neg                    # the tensor itself: it stays alive (no address reuse) until bpr()

    def bpr(self, neg):
        (_, _), (ib, ie) = self.segs
        c = self.local_bpr
        prebuilt = getattr(self, "_buckets_for", None) is neg
        self._buckets_for = None
        fn = self.L.lgcn_bpr_owner_passes if prebuilt else self.L.lgcn_bpr_owner
        self._lib.check(fn(
            byref(c), self.final.data_ptr(), self.rnorm.data_ptr(), neg.data_ptr(), self.P, self.G.data_ptr(),
            self.zg.data_ptr(), self.neg_count.data_ptr(), self.accum.data_ptr(), byref(self.bpr_ws),
            0, c.n_out_tasks, 0, c.n_in_tasks, ib - self.nu, ie - self.nu, self._p(), self._s()))

    def _bwd_call(self, gref, j, coeff, tb, te, rb, re):
        last = j == self.k
        zin = self.zg if j == 1 else self.z[j & 1]
        zout = None if last else self.z[(j - 1) & 1].data_ptr()
        reg = 2.0 * coeff / (64.0 * self.P)
        self._lib.check(self.L.lgcn_bwd_layer(
            gref, self.G.data_ptr(), j, self.k, zin.data_ptr(), zout, self.uw.data_ptr(), self.iw.data_ptr(),
            self.neg_count.data_ptr(), reg, self.grad.data_ptr() if last else None, self.accum.data_ptr(),
            tb, te, rb, re, self._p(), self._s()))

    def bwd_layer(self, j, bpr_coeff):
        if j == self.k:
            for rb, re in self.inactive_segs:
                self._bwd_call(self.g.ref, j, bpr_coeff, 0, 0, rb, re)
        self._bwd_call(byref(self.local), j, bpr_coeff, 0, self.local.n_out_tasks, 0, 0)

    def zbuf(self, j):
        return self.z[(j - 1) & 1]

    def clip_adam(self, bpr_coeff):
        for rb, re in self.segs:
            self._lib.check(self.L.lgcn_clip_adam_rows(
                byref(self.adam), self.uw.data_ptr(), self.iw.data_ptr(), self.nu, self.ni, self.grad.data_ptr(),
                self.accum.data_ptr(), self.P, bpr_coeff, self.loss.data_ptr(), rb, re, self._s()))


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
        if plan is None:
            ind, outd = ops.degrees()
            plan = ShardPlan.build(ind, outd, ops.nu, self.comm.world)
        self.plan = plan
        self.segs = plan.segments(self.comm.rank)
        self.k = ops.k
        ops.bind(plan, self.comm.rank)
        if self.comm.world > 1 and not getattr(ops, "p2p", False):
            self._gather_triplet_index()

    def _gather(self, buf: torch.Tensor, produced_by_kernel: bool = True) -> None:
        """Make every rank's copy of ``buf`` complete.  Tables produced by a p2p-enabled kernel are
        already being written into all copies; only a barrier is needed."""
        if produced_by_kernel and getattr(self.ops, "p2p", False):
            self.ops.peer_barrier()
            return
        self.comm.allgather_rows(buf, self.plan.user_ptr)
        self.comm.allgather_rows(buf, self.plan.item_ptr)

    def _gather_triplet_index(self) -> None:
        """Collective fallback: every triplet was indexed by exactly one rank (its user's owner), the other ranks hold
        zeros there."""
        o = self.ops
        if o.num_triplets:
            self.comm.allreduce(o.trip_user)
            self.comm.allreduce(o.trip_pos)

    def step(self, neg: Optional[torch.Tensor] = None, num_items: Optional[int] = None) -> torch.Tensor:
        """One step; ``neg`` is the FULL [P] negative vector (identical on every rank), or None: sampled here the way the
        reference does (uniform ``randint`` per triplet, utils/helpers.py:79-80; all ranks must share the torch CUDA RNG
        state).  Returns the loss as a device tensor (same on every rank).

        Work that does not depend on the exchanged tables -- sampling, grouping the negatives by item -- is issued
        between the kernel that produces y_0 and the barrier that waits for the peers' rows: it runs while the NVLink
        transfers of the first exchange are in flight."""
        o, k = self.ops, self.k
        p2p = getattr(o, "p2p", False)
        o.step_begin()
        o.prescale()
        if neg is None:
            ni = self.plan.num_items if num_items is None else num_items
            neg = torch.randint(0, ni, (o.num_triplets,), device=self.user_w.device)
        if hasattr(o, "bpr_buckets"):
            o.bpr_buckets(neg)
        self._gather(o.y[0])
        for layer in range(1, k + 1):
            o.fwd_layer(layer)
            if layer < k:
                self._gather(o.y[layer])
        self._gather(o.final)                 # normalised rows; 1/||final|| stays with the owner
        o.bpr(neg)
        self._gather(o.zg)
        for j in range(1, k + 1):
            o.bwd_layer(j, self.bpr_coeff)
            if j < k:
                self._gather(o.zbuf(j))
        if p2p:
            o.allreduce_accum()
        else:
            self.comm.allreduce(o.accum)
        o.clip_adam(self.bpr_coeff)
        return o.loss

    def step_sampled(self, num_items: Optional[int] = None, use_graph: bool = True) -> torch.Tensor:
        """One step with the reference's negative sampling (uniform ``randint`` per triplet,
        utils/helpers.py:79-80) done on the device.  All ranks must hold the same torch CUDA RNG state
        (same ``torch.manual_seed``) so that they draw identical negatives.  From the 4th call on the
        whole step -- sampling, kernels, peer barriers -- is replayed as ONE CUDA graph launch per rank."""
        ni = self.plan.num_items if num_items is None else num_items
        p, dev = self.ops.num_triplets, self.user_w.device
        if getattr(self, "_graph_for", None) != getattr(self.ops, "graph_version", 0):
            self.drop_graph()                          # the edge list was reloaded since the capture
            self._graph_for = getattr(self.ops, "graph_version", 0)
        self._calls = getattr(self, "_calls", 0) + 1
        if not use_graph or not self.user_w.is_cuda or getattr(self, "_graph_failed", False):
            return self.step(None, ni)
        if getattr(self, "_graph", None) is None:
            if self._calls <= 3:                       # eager warm-up (allocator, lazy module loading)
                return self.step(None, ni)
            try:
                torch.cuda.synchronize(dev)
                self.comm.barrier()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._graph_loss = self.step(None, ni)
                self._graph = g
            except Exception as exc:                   # capture not supported for some op: stay eager
                self._graph_failed, self._graph_error = True, f"{type(exc).__name__}: {exc}"
                torch.cuda.synchronize(dev)
                return self.step(None, ni)
        self._graph.replay()
        return self._graph_loss

    def drop_graph(self) -> None:
        """Forget the captured step (its launches hold the addresses of the current CSR arrays)."""
        self._graph = None
        self._calls = 0

    def gather_weights(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Full, up-to-date tables on every rank (for state_dict / best_model.pth)."""
        nu = self.plan.num_users
        full = torch.cat([self.user_w, self.item_w])
        self._gather(full, produced_by_kernel=False)
        return full[:nu].clone(), full[nu:].clone()

    def propagate_only(self) -> torch.Tensor:
        """Forward propagation alone (BASELINE config C3's per-layer timing): returns final [N,64]."""
        o, k = self.ops, self.k
        o.prescale()
        self._gather(o.y[0])
        for layer in range(1, k + 1):
            o.fwd_layer(layer, normalized=False)
            if layer < k:
                self._gather(o.y[layer])
        self._gather(o.final)
        return o.final


class HostShardPipeline:
    """Training steps whose edge list arrives from the HOST every step (what the reference's loop does with
    ``batch.to(device)``, utils/train_test.py:87): the rank's shard (pinned int64 [2,E_r] + its global triplet numbers)
    is copied on a side stream into one of two device buffers while the previous step is still building its CSR pair
    and training, so a step costs max(copy, build + train) instead of their sum.  Every step still uploads and
    rebuilds its own graph; nothing is cached between steps."""

    def __init__(self, trainer: ShardedTrainer, host_edges: torch.Tensor, host_trip: Optional[torch.Tensor] = None):
        if not host_edges.is_pinned() or (host_trip is not None and not host_trip.is_pinned()):
            raise ValueError("HostShardPipeline needs pinned host tensors (the copies must be asynchronous)")
        self.trainer, self.ops = trainer, trainer.ops
        dev = self.ops.dev
        self.host_edges, self.host_trip = host_edges, host_trip
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.edges = [torch.empty(host_edges.shape, dtype=host_edges.dtype, device=dev) for _ in range(2)]
        self.trip = [None if host_trip is None else torch.empty(host_trip.shape, dtype=host_trip.dtype, device=dev)
                     for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [None, None]                      # recorded on the main stream when a buffer's step has been built
        self.count, self.pending = 0, None
        self.bytes_per_step = host_edges.numel() * host_edges.element_size() + (
            0 if host_trip is None else host_trip.numel() * host_trip.element_size())

    def _copy(self, slot: int) -> None:
        cs = self.copy_stream
        if self.free[slot] is not None:
            cs.wait_event(self.free[slot])
        else:
            cs.wait_stream(torch.cuda.current_stream(self.ops.dev))      # the buffers were allocated on the main stream
        with torch.cuda.stream(cs):
            self.edges[slot].copy_(self.host_edges, non_blocking=True)
            if self.host_trip is not None:
                self.trip[slot].copy_(self.host_trip, non_blocking=True)
            self.ready[slot].record(cs)
        self.pending = slot

    def step(self, num_items: Optional[int] = None) -> torch.Tensor:
        slot = self.count & 1
        if self.pending != slot:                      # first call (or after a reset): nothing in flight yet
            self._copy(slot)
        main = torch.cuda.current_stream(self.ops.dev)
        main.wait_event(self.ready[slot])
        self._copy(slot ^ 1)                          # the NEXT step's upload overlaps this step's build + kernels
        self.trainer.drop_graph()
        self.ops.load_shard(self.edges[slot], self.trip[slot])
        ev = torch.cuda.Event()
        ev.record(main)                               # the CSR pair is built: the edge buffer may be overwritten
        self.free[slot] = ev
        loss = self.trainer.step_sampled(num_items, use_graph=False)
        self.count += 1
        return loss
