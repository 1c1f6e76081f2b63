"""Training / evaluation loop with the reference's function signatures
(/root/reference/utils/train_test.py:18-256) running on the sm_100a kernels.

Two ways through ``train``:
  * ``FusedAdam`` optimiser (what ``train_model`` creates): ONE C-ABI call per batch
    (``lgcn_train_step`` = forward, BPR loss, backward, clip, Adam -- utils/train_test.py:88-96),
    no per-batch host sync; the per-batch losses are read back once per epoch.
  * any ``torch.optim`` optimiser: the reference's own sequence (compute_embeddings -> bpr_loss ->
    backward -> clip_grad_norm_ -> step) with the model forward/backward and the loss on the
    kernels via autograd.
"""
from __future__ import annotations

import ctypes
from ctypes import byref
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .._lib import BatchedGraphs, CAdam, CGraph, DIM, LgcnError, StepBuffers, check, lib, require_cuda, stream_ptr
from .helpers import get_triplets_indices, sample_negative


# ----------------------------------------------------------------------------------------------
# bpr_loss on gathered rows (drop-in signature)
# ----------------------------------------------------------------------------------------------

SPARSE_STEPS = True      # train(): use the touched-rows step for small batches (set False to force dense)
CUDA_GRAPHS = True       # train(): replay each batch's step as one CUDA graph from its second epoch on
EPOCH_KERNEL = True      # train(): consecutive sparse steps run inside ONE persistent cooperative launch
STAGED_HOST_BATCHES = True   # train(): host-resident batches are uploaded and built together (K0b)
STAGE_MAX_LISTS = 1024       # per lgcn_graph_build_batched call
STAGE_MAX_CELLS = 1 << 27    # lists x (N+1)
STAGE_MAX_EDGES = 1 << 27


class _BprRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uf, u0, pf, p0, nf, n0, coeff):
        ts = [require_cuda(t, "bpr_loss input", torch.float32).contiguous() for t in (uf, u0, pf, p0, nf, n0)]
        p = ts[0].size(0)
        if any(t.shape != (p, DIM) for t in ts):
            raise LgcnError(f"bpr_loss expects six [P,{DIM}] tensors")
        dev = ts[0].device
        accum = torch.empty(2, dtype=torch.float64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        if p == 0:                       # mean over an empty set, as in the reference (NaN)
            return loss.fill_(float("nan"))
        check(lib().lgcn_bpr_rows(*[t.data_ptr() for t in ts], p, coeff, accum.data_ptr(), loss.data_ptr(), None,
                                  None, None, None, None, None, None, stream_ptr(dev)))
        ctx.save_for_backward(*ts)
        ctx.coeff = coeff
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        ts = ctx.saved_tensors
        p, dev = ts[0].size(0), ts[0].device
        gs = grad_out.to(torch.float32).contiguous()
        grads = [torch.empty_like(t) for t in ts]
        # order of the C signature: g_uf, g_u0, g_pf, g_p0, g_nf, g_n0  == order of the inputs
        check(lib().lgcn_bpr_rows(*[t.data_ptr() for t in ts], p, ctx.coeff, None, None, gs.data_ptr(),
                                  *[g.data_ptr() for g in grads], stream_ptr(dev)))
        return (*grads, None)


def bpr_loss(emb_users_final: torch.Tensor, emb_users: torch.Tensor,
             emb_pos_items_final: torch.Tensor, emb_pos_items: torch.Tensor,
             emb_neg_items_final: torch.Tensor, emb_neg_items: torch.Tensor,
             bpr_coeff: float = 5e-3) -> torch.Tensor:
    """-mean(softplus(10 (cos+ - cos-)))/10 + bpr_coeff * mean(u0^2 + p0^2 + n0^2)
    (utils/train_test.py:18-51), one fused kernel, differentiable w.r.t. all six inputs."""
    return _BprRows.apply(emb_users_final, emb_users, emb_pos_items_final, emb_pos_items,
                          emb_neg_items_final, emb_neg_items, float(bpr_coeff))


def normalize_embedding(emb: torch.Tensor) -> torch.Tensor:
    """Row-wise L2 normalisation without epsilon (utils/train_test.py:53-64)."""
    return emb / torch.norm(emb, p=2, dim=1, keepdim=True)


# ----------------------------------------------------------------------------------------------
# fused optimiser state
# ----------------------------------------------------------------------------------------------

class FusedAdam:
    """Adam(lr, betas=(0.9, 0.999), eps=1e-8) over both embedding tables preceded by
    clip_grad_norm_(max_norm) -- utils/train_test.py:95-96,236 -- as state for the fused step.
    exp_avg / exp_avg_sq are flat [N,64] tensors; the step counter lives on the device so that a
    training step needs no host round trip."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0):
        w = model.user_embedding.weight
        require_cuda(w, "model parameters", torch.float32)
        self.model = model
        n = model.num_users + model.num_items
        self.exp_avg = torch.zeros(n, DIM, dtype=torch.float32, device=w.device)
        self.exp_avg_sq = torch.zeros(n, DIM, dtype=torch.float32, device=w.device)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=w.device)
        self.row_step = torch.zeros(n, dtype=torch.int32, device=w.device)   # sparse steps: per-row progress
        self.pending = False          # True while some rows still owe zero-gradient updates (sparse steps)
        self.dirty = False            # True when a dense step left dL/dfinal / the negative histogram non-zero
        self.graphs = {}              # id(Graph) -> (Graph, CUDAGraph | None, loss slot | None)
        self.captured = 0
        self.graph_generation = 0
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self._bc_table(1 << 18)
        self.buffers = StepBuffers(n, model.num_items, model.num_layers, w.device)
        self.losses = torch.zeros(8192, dtype=torch.float32, device=w.device)   # per-batch losses (see train)
        self.stage = _Stage()         # device/pinned buffers of the host-batch pipeline (see _run_staged)
        self.c = CAdam()
        self._fill()

    def _bc_table(self, length: int):
        """(lr/(1-beta1^t), sqrt(1-beta2^t)) for t < length, in double like torch's _single_tensor_adam,
        cast once to fp32.  Extended on demand (host_steps tracks how far training may have got)."""
        t = np.arange(length, dtype=np.float64)
        with np.errstate(divide="ignore"):
            ss = self.lr / (1.0 - np.power(self.betas[0], t))
        bc2 = np.sqrt(1.0 - np.power(self.betas[1], t))
        tab = np.stack([ss, bc2], axis=1).astype(np.float32)
        tab[0] = 0.0
        self.bc_table = torch.from_numpy(tab).to(self.exp_avg.device).contiguous()
        self.host_steps = getattr(self, "host_steps", 0)

    def _fill(self):
        c = self.c
        c.lr, c.beta1, c.beta2, c.eps, c.max_norm = self.lr, self.betas[0], self.betas[1], self.eps, self.max_norm
        c.step, c.m, c.v = self.step_count.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        c.bc_table, c.bc_len = self.bc_table.data_ptr(), self.bc_table.shape[0]
        c.row_step = self.row_step.data_ptr()

    def _count_step(self, n: int = 1):
        self.host_steps += n
        if self.host_steps + 2 >= self.bc_table.shape[0]:
            self._bc_table(2 * max(self.bc_table.shape[0], self.host_steps + 2))
            self._fill()
            self.graphs.clear()           # captured launches hold the old table address
            self.captured = 0

    def flush(self):
        """Replay the zero-gradient updates that sparse steps deferred, so that every row of the
        weights / moments is at the current step (called before anything else reads the weights)."""
        if self.pending:
            m = self.model
            check(lib().lgcn_adam_flush(byref(self.c), m.user_embedding.weight.data_ptr(),
                                        m.item_embedding.weight.data_ptr(), m.num_users, m.num_items,
                                        stream_ptr(self.exp_avg.device)))
            self.pending = False

    def zero_grad(self):
        pass

    def zero_sparse_invariants(self):
        """The touched-rows steps need dL/dfinal and the negative histogram all-zero on entry (they leave them
        all-zero on exit); a dense step in between leaves both dirty."""
        if self.dirty:
            self.buffers.grad_final.zero_()
            self.buffers.neg_count.zero_()
            self.dirty = False

    def graph_ready(self, g, sparse: bool) -> bool:
        """Capturing freezes buffer addresses: the scratch must already fit this batch, the bias
        table must have room, and the zero-invariants of the sparse step must hold."""
        return (self.buffers.trip_scratch.numel() >= 2 * g.num_triplets
                and self.host_steps + (1 << 12) < self.bc_table.shape[0])

    def state_dict(self):
        self.flush()
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_count,
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "max_norm": self.max_norm}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.step_count.copy_(sd["step"])
        self.lr, self.betas, self.eps, self.max_norm = sd["lr"], tuple(sd["betas"]), sd["eps"], sd["max_norm"]
        self.host_steps = int(self.step_count)
        self.row_step.fill_(self.host_steps)
        self.pending = False
        self.graphs.clear()               # captured launches hold the old bias table and the old hyper-parameters by value
        self.captured = 0
        self.dirty = True                 # whatever the scratch holds belongs to another timeline
        self.buffers.neg_flag.zero_()     # step stamps of another timeline
        self.buffers.act_stamp.zero_()
        self._bc_table(max(1 << 16, 2 * self.host_steps + 4))
        self._fill()


class _Stage:
    """Buffers the host-batch pipeline reuses from epoch to epoch (no per-epoch cudaMalloc)."""

    def __init__(self):
        self.edges = self.arena = self.workspace = self.current = self.uploading = None


def _check_neg(neg: torch.Tensor, num_items: int) -> None:
    """Caller-supplied negatives index rows of the item table on the device: refuse ids outside [0, I) here (one
    small reduction + read-back; negatives sampled by this module need no check)."""
    if neg.numel() and bool(((neg < 0) | (neg >= num_items)).any()):
        raise LgcnError(f"neg holds item ids outside [0, {num_items})")


def train_step(model, optimizer: FusedAdam, edge_index: torch.Tensor, neg: Optional[torch.Tensor] = None,
               loss_out: Optional[torch.Tensor] = None, bpr_coeff: float = 5e-3, sparse: bool = False) -> torch.Tensor:
    """The loop body utils/train_test.py:88-96 for one batch as one C-ABI call.  ``neg`` defaults to
    the reference's sampling (utils/helpers.py:79-80).  Returns a 0-dim DEVICE tensor (no sync).

    ``sparse=True`` runs ``lgcn_train_step_sparse``: same arithmetic, but only the rows the batch
    touches are visited; the other rows' zero-gradient Adam updates stay pending until they are next
    touched or ``optimizer.flush()`` is called (``train`` does that at the end of the epoch)."""
    g = model.graph(edge_index)
    dev = edge_index.device
    if g.num_triplets == 0:
        raise LgcnError("batch has no user->movie edge: the reference's loss is NaN here (SURVEY App. B #13)")
    if neg is None:
        neg = torch.randint(0, model.num_items, (g.num_triplets,), device=dev)
    else:
        require_cuda(neg, "neg", torch.int64)
        _check_neg(neg, model.num_items)
    if neg.numel() != g.num_triplets:
        raise LgcnError(f"neg has {neg.numel()} entries, the batch has {g.num_triplets} user->movie edges")
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=dev)
    uw, iw = model.user_embedding.weight, model.item_embedding.weight
    if not (uw.is_contiguous() and iw.is_contiguous()):
        raise LgcnError("embedding weights must be contiguous")
    optimizer.buffers.ensure_triplets(g.num_triplets)
    optimizer._count_step()
    if sparse:
        optimizer.zero_sparse_invariants()
        optimizer.pending = True
    else:
        optimizer.flush()
    _launch_step(model, optimizer, g, neg.contiguous(), loss_out, bpr_coeff, sparse)
    return loss_out


def _launch_step(model, optimizer, g, neg, loss_out, bpr_coeff, sparse):
    _launch_step_raw(model, optimizer, g, neg.data_ptr(), loss_out.data_ptr(), bpr_coeff, sparse, stream_ptr(neg.device))


def _launch_step_raw(model, optimizer, g, neg_ptr, loss_ptr, bpr_coeff, sparse, stream):
    uw, iw = model.user_embedding.weight, model.item_embedding.weight
    fn = lib().lgcn_train_step_sparse if sparse else lib().lgcn_train_step
    if not sparse:
        optimizer.dirty = True
    check(fn(g.ref, uw.data_ptr(), iw.data_ptr(), model.num_layers, neg_ptr, bpr_coeff, byref(optimizer.c),
             optimizer.buffers.ref, loss_ptr, stream))


class _SparseRun:
    """Consecutive sparse steps waiting to be launched as ONE lgcn_train_steps_sparse call."""

    def __init__(self, model, optimizer, device, weights, slots, cap):
        self.model, self.opt, self.device = model, optimizer, device
        self.weights, self.slots, self.cap = weights, slots, cap
        self.graphs, self.edges = [], []
        self.cg = None                    # optional: a ready-made contiguous CGraph array covering exactly the run

    def add(self, g, num_edges: int):
        self.graphs.append(g)
        self.edges.append(num_edges)

    def flush(self):
        if not self.graphs:
            return
        b = len(self.graphs)
        slot0 = len(self.weights)
        if slot0 + b > self.cap:
            raise LgcnError(f"more than {self.cap} batches in one epoch: enlarge FusedAdam.losses")
        _launch_steps(self.model, self.opt, self.graphs, None, self.opt.losses.data_ptr() + 4 * slot0, 5e-3, self.device,
                      self.cg if self.cg is not None and len(self.cg) == b else None)
        self.weights.extend(self.edges)
        self.slots.extend(range(slot0, slot0 + b))
        self.graphs, self.edges = [], []


def _launch_steps(model, opt, graphs, neg_all, loss_ptr, bpr_coeff, device, cg=None) -> None:
    """One lgcn_train_steps_sparse call for ``graphs`` (all with triplets); ``neg_all``: the steps'
    negatives back to back, sampled here when None."""
    b = len(graphs)
    trip = [g.num_triplets for g in graphs]
    if min(trip) <= 0:
        raise LgcnError("a batch has no user->movie edge: the reference's loss is NaN here (SURVEY App. B #13)")
    if neg_all is None:
        neg_all = torch.randint(0, model.num_items, (sum(trip),), device=device)
    require_cuda(neg_all, "neg", torch.int64)
    if neg_all.numel() != sum(trip):
        raise LgcnError(f"neg has {neg_all.numel()} entries, the batches have {sum(trip)} user->movie edges")
    buf = opt.buffers
    buf.ensure_triplets(max(trip))
    L = lib()
    need = L.lgcn_train_steps_workspace_bytes(b)
    if buf.steps_ws.numel() < need:
        buf.steps_ws = torch.empty(2 * need, dtype=torch.uint8, device=device)
    opt.zero_sparse_invariants()
    opt.pending = True
    opt._count_step(b)
    if cg is None:
        cg = (CGraph * b)(*[g.c for g in graphs])
    uw, iw = model.user_embedding.weight, model.item_embedding.weight
    if not (uw.is_contiguous() and iw.is_contiguous()):
        raise LgcnError("embedding weights must be contiguous")
    check(L.lgcn_train_steps_sparse(cg, b, uw.data_ptr(), iw.data_ptr(), model.num_layers, neg_all.contiguous().data_ptr(),
                                    bpr_coeff, byref(opt.c), buf.ref, loss_ptr, buf.steps_ws.data_ptr(),
                                    buf.steps_ws.numel(), stream_ptr(device)))


def train_steps(model, optimizer: FusedAdam, edge_indices: Sequence[torch.Tensor],
                negs: Optional[Sequence[torch.Tensor]] = None, bpr_coeff: float = 5e-3) -> torch.Tensor:
    """``train_step(..., sparse=True)`` for a whole sequence of batches in ONE persistent launch
    (``lgcn_train_steps_sparse``): the loop of utils/train_test.py:86-96 over ``edge_indices`` in order.
    Returns the per-batch losses as a DEVICE tensor (no sync).  Rows the batches do not touch keep their
    zero-gradient Adam updates pending (``optimizer.flush()``)."""
    graphs = [model.graph(ei) for ei in edge_indices]
    dev = edge_indices[0].device
    neg_all = None if negs is None else torch.cat([n.reshape(-1) for n in negs])
    if negs is not None:
        _check_neg(neg_all, model.num_items)
        for g, n in zip(graphs, negs):
            if n.numel() != g.num_triplets:
                raise LgcnError(f"neg has {n.numel()} entries, the batch has {g.num_triplets} user->movie edges")
    losses = torch.empty(len(graphs), dtype=torch.float32, device=dev)
    _launch_steps(model, optimizer, graphs, neg_all, losses.data_ptr(), bpr_coeff, dev)
    return losses


def _run_staged(model, optimizer: "FusedAdam", eis, device, weights, slots, cap) -> None:
    """The loop body of utils/train_test.py:86-101 for a run of HOST-resident batches:
    ``batch.to(device)`` becomes one upload call for all the edge lists (``lgcn_upload_lists``), the per-batch
    normalisation / CSR becomes ONE batched build (K0b),
    the negatives of a run of steps come from one ``randint`` (utils/helpers.py:79-80: uniform, no
    rejection; neg k pairs with the k-th user->movie edge), then the steps run in loader order -- sparse
    ones inside one persistent launch, dense ones one call each."""
    st = optimizer.stage
    sizes = [int(ei.shape[1]) for ei in eis]
    off = np.zeros(len(eis) + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    tot = int(off[-1])
    if st.edges is None or st.edges.numel() < 2 * tot or st.edges.device != device:
        st.edges = torch.empty(int(2.5 * tot) + 64, dtype=torch.int64, device=device)
    # one C call issues the B copies (asynchronous from pinned tensors, driver-staged from pageable ones)
    keep = [ei if ei.is_contiguous() else ei.contiguous() for ei in eis]
    ptrs = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
    check(lib().lgcn_upload_lists(ptrs, off.ctypes.data, len(keep), st.edges.data_ptr(), stream_ptr(device)))
    st.uploading = keep                  # the host tensors must outlive the asynchronous copies
    bg = BatchedGraphs(st.edges, off, model.num_users, model.num_items, st.arena, st.workspace)
    st.arena, st.workspace, st.current = bg.arena, bg.workspace, bg
    run = _SparseRun(model, optimizer, device, weights, slots, cap)
    whole = True                          # every list joins one run: the build's own struct array serves as is
    for g, e in zip(bg.graphs, sizes):
        if g.num_triplets == 0:
            whole = False
            continue                      # the reference would produce NaN here (App. B #13)
        if EPOCH_KERNEL and sparse_step_pays(g):
            run.add(g, e)
            continue
        whole = False
        run.flush()
        _single_step(model, optimizer, g, e, device, weights, slots, cap)
    if whole:
        run.cg = bg._cg
    run.flush()


def _single_step(model, optimizer, g, num_edges, device, weights, slots, cap) -> None:
    """One eager fused step (sparse or dense) on an already built graph, negatives sampled here."""
    slot = len(weights)
    if slot >= cap:
        raise LgcnError(f"more than {cap} batches in one epoch: enlarge FusedAdam.losses")
    sparse = sparse_step_pays(g)
    neg = torch.randint(0, model.num_items, (g.num_triplets,), device=device)
    optimizer.buffers.ensure_triplets(g.num_triplets)
    optimizer._count_step()
    if sparse:
        optimizer.zero_sparse_invariants()
        optimizer.pending = True
    else:
        optimizer.flush()
    _launch_step(model, optimizer, g, neg, optimizer.losses[slot:slot + 1], 5e-3, sparse)
    weights.append(num_edges)
    slots.append(slot)


def sparse_step_pays(g) -> bool:
    """Touched rows (nodes with edges + at most min(P, I) distinct negatives) under half the table."""
    return SPARSE_STEPS and 2 * (g.num_active + min(g.num_triplets, g.num_items)) < g.num_nodes


def _train_epoch_fused(model, optimizer: "FusedAdam", train_loader, device) -> float:
    """One epoch on the fused path.  Batches that touch a small part of the table (Cluster-GCN) are
    collected, in loader order, into runs that execute inside ONE persistent cooperative launch
    (``lgcn_train_steps_sparse``); host-resident batches are first uploaded and built together
    (``_run_staged``).  A dense batch is one C-ABI call, replayed as a CUDA graph from the second time its
    tensor is seen (the reference's loader hands back the same tensors every epoch,
    data/dataset_handler.py:277-285).  The only host<->device sync is the loss read-back at the end."""
    device = torch.device(device)
    weights, slots = [], []
    cap = optimizer.losses.numel() // 2          # [0,cap): this epoch's eager steps; [cap,2cap): captured graphs
    if len(optimizer.graphs) > cap:              # fresh tensors every epoch: forget the never-captured ones
        optimizer.graphs = {k: v for k, v in optimizer.graphs.items() if v[1] is not None}
    pending, pending_edges = [], 0
    run = _SparseRun(model, optimizer, device, weights, slots, cap)

    def run_pending():
        nonlocal pending, pending_edges
        if pending:
            run.flush()
            _run_staged(model, optimizer, pending, device, weights, slots, cap)
            pending, pending_edges = [], 0

    for batch in train_loader:
        ei = batch.edge_index
        if STAGED_HOST_BATCHES and not ei.is_cuda and device.type == "cuda":
            if ei.shape[1] == 0:
                continue
            if ei.dtype != torch.int64 or ei.dim() != 2 or ei.size(0) != 2:
                raise LgcnError(f"edge_index must be an int64 [2,E] tensor, got {ei.dtype} {tuple(ei.shape)}")
            if (len(pending) + 1 > STAGE_MAX_LISTS or (len(pending) + 1) * (model.num_users + model.num_items + 1)
                    > STAGE_MAX_CELLS or pending_edges + ei.shape[1] > STAGE_MAX_EDGES):
                run_pending()
            pending.append(ei)
            pending_edges += ei.shape[1]
            continue
        run_pending()                     # keep the loader's order of optimiser steps
        batch = batch.to(device)
        ei = batch.edge_index
        if ei.shape[1] == 0:
            continue
        g = model.graph(ei)
        if g.num_triplets == 0:
            continue                      # the reference would produce NaN here (App. B #13)
        if optimizer.buffers.generation != optimizer.graph_generation:
            optimizer.graphs.clear()      # a scratch buffer moved: captured launches are stale
            optimizer.graph_generation, optimizer.captured = optimizer.buffers.generation, 0
        # Cluster-GCN batches touch a small part of the table: visit only those rows
        sparse = sparse_step_pays(g)
        if sparse and EPOCH_KERNEL:
            run.add(g, ei.shape[1])       # launched together with its neighbours in loader order
            continue
        run.flush()
        entry = optimizer.graphs.get(id(g)) if CUDA_GRAPHS else None
        if entry is not None and entry[0] is not g:
            entry = None
        if entry is not None and entry[1] is not None:                      # replay
            optimizer._count_step()
            if sparse:
                optimizer.zero_sparse_invariants()      # a dense step in between leaves dL/dfinal / the histogram dirty
                optimizer.pending = True
            else:
                optimizer.flush()
                optimizer.dirty = True
            entry[1].replay()
            slot = entry[2]
        elif (entry is not None and optimizer.graph_ready(g, sparse) and optimizer.captured < cap
              and len(weights) < cap):                                      # second sighting: capture
            slot = cap + optimizer.captured
            optimizer.captured += 1
            optimizer._count_step()
            if not sparse:
                optimizer.flush()
            if sparse:
                optimizer.zero_sparse_invariants()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                neg = torch.randint(0, model.num_items, (g.num_triplets,), device=device)
                _launch_step(model, optimizer, g, neg, optimizer.losses[slot:slot + 1], 5e-3, sparse)
            optimizer.graphs[id(g)] = (g, cg, slot)
            if sparse:
                optimizer.pending = True
            cg.replay()
        else:                                                               # first sighting: eager
            slot = len(weights)
            if slot >= cap:
                raise LgcnError(f"more than {cap} batches in one epoch: enlarge FusedAdam.losses")
            train_step(model, optimizer, ei, loss_out=optimizer.losses[slot:slot + 1], sparse=sparse)
            if CUDA_GRAPHS and entry is None:
                optimizer.graphs[id(g)] = (g, None, None)
        weights.append(ei.shape[1])
        slots.append(slot)
    run_pending()
    run.flush()
    optimizer.flush()
    if not weights:
        return float("nan")
    w = torch.tensor(weights, dtype=torch.float64)
    if slots == list(range(slots[0], slots[0] + len(slots))):
        losses = optimizer.losses[slots[0]: slots[0] + len(slots)].cpu().double()                    # the only sync
    else:
        losses = optimizer.losses[torch.tensor(slots, device=optimizer.losses.device)].cpu().double()
    return float((losses * w).sum() / w.sum())


def train(model: torch.nn.Module, optimizer, train_loader, device: torch.device) -> float:
    """One epoch (utils/train_test.py:66-103): edge-count-weighted mean of the batch losses."""
    model.train()
    if isinstance(optimizer, FusedAdam):
        return _train_epoch_fused(model, optimizer, train_loader, device)

    total_loss, total_w = 0.0, 0
    for batch in train_loader:
        batch = batch.to(device)
        optimizer.zero_grad()
        embs = compute_embeddings(model, batch, device)
        train_loss = bpr_loss(*embs)
        train_loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1)
        optimizer.step()
        w = batch.edge_index.shape[1]
        total_w += w
        total_loss += train_loss.item() * w
    return total_loss / total_w


def compute_embeddings(model: torch.nn.Module, data, device: torch.device) -> Tuple[torch.Tensor, ...]:
    """Final/initial rows for (user, pos, neg) of every user->movie edge (utils/train_test.py:105-134)."""
    final_user, final_item = model(data.edge_index)
    init_user, init_item = model.user_embedding.weight, model.item_embedding.weight
    u, p, n = get_triplets_indices(data.edge_index, model.num_users, model.num_items, device)
    return final_user[u], init_user[u], final_item[p], init_item[p], final_item[n], init_item[n]


def eval_loss(model, edge_index: torch.Tensor, neg: torch.Tensor, buffers: Optional[StepBuffers] = None,
              bpr_coeff: float = 5e-3) -> torch.Tensor:
    """Loss of evaluate() without materialising the six gathers (utils/train_test.py:153-156)."""
    g = model.graph(edge_index)
    dev = edge_index.device
    if buffers is None:
        buffers = StepBuffers(g.num_nodes, model.num_items, model.num_layers, dev)
    require_cuda(neg, "neg", torch.int64)
    _check_neg(neg, model.num_items)
    if neg.numel() != g.num_triplets:
        raise LgcnError(f"neg has {neg.numel()} entries, the edge list has {g.num_triplets} user->movie edges")
    out = torch.empty(1, dtype=torch.float32, device=dev)
    check(lib().lgcn_eval_loss(g.ref, model.user_embedding.weight.data_ptr(), model.item_embedding.weight.data_ptr(),
                               model.num_layers, neg.contiguous().data_ptr(), bpr_coeff, buffers.ref, out.data_ptr(),
                               stream_ptr(dev)))
    return out


def evaluate(model: torch.nn.Module, test_data, device: torch.device, top_k: int = 100) -> Tuple[float, float]:
    """Loss over the given edges + the reference's sampled recall on LAYER-0 rows
    (utils/train_test.py:136-163)."""
    model.eval()
    with torch.no_grad():
        test_data = test_data.to(device)
        u, p, n = get_triplets_indices(test_data.edge_index, model.num_users, model.num_items, device)
        test_loss = eval_loss(model, test_data.edge_index, n).item()
        uw, iw = model.user_embedding.weight, model.item_embedding.weight
        recall_at_k = compute_recall_at_k((uw[u], iw[p], iw[n]), k=top_k)
    return test_loss, recall_at_k


def compute_recall_at_k(embs, k: int = 20, num_samples: int = 10, sample_size: int = 100,
                        sampled: Optional[Sequence[np.ndarray]] = None) -> float:
    """The reference's sampled Recall@k (utils/train_test.py:165-212), definition preserved: the
    candidate pool is the [pos; neg] rows, a hit is any top-k column < P, the denominator is P.
    Scoring + top-k run in the fused kernel (no 100 x 2P score matrix).  ``sampled`` lets a test
    supply the np.random.choice draws (:187)."""
    from .recommend import score_topk
    user_embs, pos_item_embs, neg_item_embs = embs
    p = pos_item_embs.size(0)
    cand = torch.cat((pos_item_embs, neg_item_embs)).contiguous()
    num_users = user_embs.size(0)
    total = 0.0
    for s in range(num_samples):
        idx = sampled[s] if sampled is not None else np.random.choice(num_users, sample_size, replace=False)
        rows = user_embs[torch.as_tensor(idx, device=user_embs.device)].contiguous()
        top_idx, _ = score_topk(rows, cand, k, normalize=True)
        hits = ((top_idx >= 0) & (top_idx < p)).sum(dim=1).to(torch.float32)   # -1 pads lists shorter than k
        total += (hits / p).mean().item()
    return total / num_samples


def train_model(model: torch.nn.Module, train_loader, val_data, test_data, device: torch.device,
                epochs: int = 1, lr: float = 0.001):
    """utils/train_test.py:214-256 (Adam lr, best-val-recall checkpoint to best_model.pth)."""
    hist_train_loss, hist_val_loss, hist_val_recall = [], [], []
    optimizer = FusedAdam(model, lr=lr)
    best_recall = 0
    for epoch in range(epochs):
        loss = train(model, optimizer, train_loader, device)
        val_loss, recall_at_k = evaluate(model, val_data, device)
        hist_train_loss.append(loss)
        hist_val_loss.append(val_loss)
        hist_val_recall.append(recall_at_k)
        print(f"Epoch: {epoch:03d}, Train Loss: {loss:.4f}, Val Loss: {val_loss:.4f}, "
              f"Recall@k: {recall_at_k:.6f}, k=100")
        if recall_at_k > best_recall:
            best_recall = recall_at_k
            torch.save(model.state_dict(), "best_model.pth")
    test_loss, recall_at_k = evaluate(model, test_data, device)
    print(f"Test Loss: {test_loss:.4f}, Recall@k: {recall_at_k:.6f}, k=100")
    return model, hist_train_loss, hist_val_loss, hist_val_recall
