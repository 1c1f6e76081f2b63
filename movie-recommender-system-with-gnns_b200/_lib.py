"""ctypes binding of csrc/liblgcn_b200.so (the C ABI declared in include/lgcn_b200.h).

PyTorch is used here only for device memory and streams: every array the library reads or writes
is a torch tensor allocated by this module, passed as a raw pointer.  There is NO fallback: if the
shared library is missing, or a tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64,
                    c_size_t, c_uint8, c_void_p)
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGCN_LIB_PATH") or os.path.join(_HERE, "csrc", "liblgcn_b200.so")   # env: tuning variants

DIM = 64
ROW_SPLIT = 512
PARTIAL_STRIDE = 80

EXPORTS = [
    "lgcn_last_error", "lgcn_version", "lgcn_graph_sizes_query", "lgcn_graph_build",
    "lgcn_propagate_fwd", "lgcn_propagate_bwd", "lgcn_bpr_fwd_bwd", "lgcn_step_begin",
    "lgcn_clip_adam", "lgcn_train_step", "lgcn_eval_loss", "lgcn_partition_metis",
    "lgcn_cluster_extract_workspace_bytes", "lgcn_cluster_extract", "lgcn_score_topk",
    "lgcn_spmm", "lgcn_bpr_rows", "lgcn_prescale", "lgcn_fwd_layer", "lgcn_bwd_layer",
    "lgcn_bpr_fwd_bwd_range", "lgcn_clip_adam_rows", "lgcn_train_step_sparse", "lgcn_adam_flush",
    "lgcn_peer_barrier", "lgcn_score_topk_ex", "lgcn_graph_batched_sizes", "lgcn_graph_build_batched",
    "lgcn_train_steps_workspace_bytes", "lgcn_train_steps_sparse", "lgcn_probe_gather",
    "lgcn_score_topk_workspace_bytes", "lgcn_upload_lists",
    "lgcn_bpr_owner", "lgcn_fwd_layer_ex", "lgcn_triplet_index", "lgcn_graph_remap_triplets", "lgcn_peer_allreduce4",
    "lgcn_to_undirected_workspace_bytes", "lgcn_to_undirected", "lgcn_bpr_buckets", "lgcn_bpr_owner_passes",
    "lgcn_label_vote",
]


class LgcnError(RuntimeError):
    pass


class CTask(Structure):
    _fields_ = [(n, c_int32) for n in ("row", "begin", "end", "slot", "part", "nparts", "deg_in", "deg_out")]


class CGraph(Structure):
    _fields_ = [
        ("num_nodes", c_int32), ("num_users", c_int32), ("num_edges", c_int64), ("num_triplets", c_int64),
        ("in_ptr", c_void_p), ("in_nbr", c_void_p), ("in_trip", c_void_p),
        ("out_ptr", c_void_p), ("out_nbr", c_void_p), ("out_trip", c_void_p),
        ("dis", c_void_p), ("active", c_void_p), ("in_tasks", c_void_p), ("out_tasks", c_void_p),
        ("n_in_tasks", c_int32), ("n_out_tasks", c_int32),
        ("n_in_user_tasks", c_int32), ("n_out_user_tasks", c_int32),
        ("n_in_slots", c_int32), ("n_out_slots", c_int32),
        ("partials", c_void_p), ("slot_counters", c_void_p),
        ("num_active", c_int32), ("row_split", c_int32), ("active_list", c_void_p),
        ("sched", c_void_p), ("in_src_sorted", c_int32), ("reserved0", c_int32),
    ]


class CGraphSizes(Structure):
    _fields_ = [(n, c_size_t) for n in ("ptr_bytes", "nbr_bytes", "dis_bytes", "active_bytes", "task_bytes",
                                        "partial_bytes", "counter_bytes", "workspace_bytes", "active_list_bytes")]


class CBatchedSizes(Structure):
    _fields_ = [("arena_bytes", c_size_t), ("workspace_bytes", c_size_t)]


class CAdam(Structure):
    _fields_ = [("lr", c_double), ("beta1", c_double), ("beta2", c_double), ("eps", c_double),
                ("max_norm", c_double), ("step", c_void_p), ("m", c_void_p), ("v", c_void_p),
                ("bc_table", c_void_p), ("bc_len", c_int64), ("row_step", c_void_p)]


class CPeers(Structure):
    _fields_ = [("world", c_int32), ("rank", c_int32), ("mc_base", c_void_p), ("base", c_void_p * 8)]


class CBprOwnerWs(Structure):
    _fields_ = [("trip_user", c_void_p), ("trip_pos", c_void_p), ("bucket_ptr", c_void_p), ("bucket_cursor", c_void_p),
                ("bucket", c_void_p), ("bucket_cap", c_int64), ("scalars", c_void_p), ("sched", c_void_p)]


class CStepBuffers(Structure):
    _fields_ = [("final_emb", c_void_p), ("rnorm", c_void_p), ("grad_final", c_void_p), ("grad_e0", c_void_p),
                ("work", c_void_p), ("work_bytes", c_size_t), ("neg_count", c_void_p),
                ("trip_scratch", c_void_p), ("accum", c_void_p),
                ("neg_flag", c_void_p), ("neg_list", c_void_p), ("neg_list_count", c_void_p),
                ("act_stamp", c_void_p)]


_lib = None


def lib():
    """The loaded shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LgcnError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'` -- this package has no CPU or PyTorch fallback")
    L = ctypes.CDLL(LIB_PATH)
    L.lgcn_last_error.restype = c_char_p
    L.lgcn_version.restype = c_int
    L.lgcn_graph_sizes_query.argtypes = [c_int64, c_int64, POINTER(CGraphSizes)]
    L.lgcn_graph_build.argtypes = [c_void_p, c_int64, c_int64, c_int64, POINTER(CGraph), c_void_p, c_size_t, c_void_p]
    L.lgcn_propagate_fwd.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]
    L.lgcn_propagate_bwd.argtypes = [POINTER(CGraph), c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float,
                                     c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.lgcn_bpr_fwd_bwd.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]
    L.lgcn_step_begin.argtypes = [POINTER(CAdam), c_void_p, c_void_p]
    L.lgcn_clip_adam.argtypes = [POINTER(CAdam), c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                 c_float, c_void_p, c_void_p]
    L.lgcn_train_step.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_int, c_void_p, c_float, POINTER(CAdam),
                                  POINTER(CStepBuffers), c_void_p, c_void_p]
    L.lgcn_eval_loss.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_int, c_void_p, c_float,
                                 POINTER(CStepBuffers), c_void_p, c_void_p]
    L.lgcn_partition_metis.argtypes = [c_int64, c_void_p, c_void_p, c_int64, c_void_p]
    L.lgcn_cluster_extract_workspace_bytes.argtypes = [c_int64, c_int64, c_int64]
    L.lgcn_cluster_extract_workspace_bytes.restype = c_size_t
    L.lgcn_cluster_extract.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_size_t, c_void_p]
    L.lgcn_label_vote.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.lgcn_to_undirected_workspace_bytes.argtypes = [c_int64]
    L.lgcn_to_undirected_workspace_bytes.restype = c_size_t
    L.lgcn_to_undirected.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    L.lgcn_score_topk.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_void_p]
    L.lgcn_score_topk_ex.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int,
                                     c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]
    L.lgcn_score_topk_workspace_bytes.argtypes = [c_int64]
    L.lgcn_score_topk_workspace_bytes.restype = c_size_t
    L.lgcn_spmm.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_int, c_void_p]
    L.lgcn_bpr_rows.argtypes = [c_void_p] * 6 + [c_int64, c_float, c_void_p, c_void_p, c_void_p] + [c_void_p] * 6 + [c_void_p]
    L.lgcn_prescale.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_int64, c_int64, c_void_p, POINTER(CPeers), c_void_p]
    L.lgcn_fwd_layer.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 7 + \
        [c_int, c_int, c_int64, c_int64, POINTER(CPeers), c_void_p]
    L.lgcn_fwd_layer_ex.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 7 + \
        [c_int, c_int, c_int64, c_int64, c_int, POINTER(CPeers), c_void_p]
    L.lgcn_bwd_layer.argtypes = [POINTER(CGraph), c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, POINTER(CPeers),
                                 c_void_p]
    L.lgcn_bpr_fwd_bwd_range.argtypes = [POINTER(CGraph)] + [c_void_p] * 7 + [c_int, c_int, c_int64, c_int64, c_void_p]
    L.lgcn_clip_adam_rows.argtypes = [POINTER(CAdam), c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                      c_float, c_void_p, c_int64, c_int64, c_void_p]
    L.lgcn_peer_barrier.argtypes = [POINTER(CPeers), c_void_p, c_void_p, c_void_p]
    L.lgcn_peer_allreduce4.argtypes = [POINTER(CPeers), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.lgcn_bpr_owner.argtypes = [POINTER(CGraph), c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_void_p, POINTER(CBprOwnerWs), c_int, c_int, c_int, c_int, c_int64, c_int64,
                                 POINTER(CPeers), c_void_p]
    L.lgcn_bpr_owner_passes.argtypes = L.lgcn_bpr_owner.argtypes
    L.lgcn_bpr_buckets.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, POINTER(CBprOwnerWs), c_void_p]
    L.lgcn_triplet_index.argtypes = [POINTER(CGraph), c_int, c_int, c_void_p, c_void_p, POINTER(CPeers), c_void_p]
    L.lgcn_graph_remap_triplets.argtypes = [POINTER(CGraph), c_void_p, c_void_p]
    L.lgcn_train_step_sparse.argtypes = L.lgcn_train_step.argtypes
    L.lgcn_adam_flush.argtypes = [POINTER(CAdam), c_void_p, c_void_p, c_int64, c_int64, c_void_p]
    L.lgcn_graph_batched_sizes.argtypes = [c_int64, c_int64, c_void_p, POINTER(CBatchedSizes)]
    L.lgcn_graph_build_batched.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t,
                                           c_void_p, c_size_t, c_void_p]
    L.lgcn_upload_lists.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    L.lgcn_probe_gather.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]
    L.lgcn_train_steps_workspace_bytes.argtypes = [c_int64]
    L.lgcn_train_steps_workspace_bytes.restype = c_size_t
    L.lgcn_train_steps_sparse.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_float, POINTER(CAdam),
                                          POINTER(CStepBuffers), c_void_p, c_void_p, c_size_t, c_void_p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("lgcn_last_error", "lgcn_cluster_extract_workspace_bytes", "lgcn_train_steps_workspace_bytes",
                        "lgcn_score_topk_workspace_bytes", "lgcn_to_undirected_workspace_bytes"):
            fn.restype = c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise LgcnError(f"lgcn error {rc}: {lib().lgcn_last_error().decode()}")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not t.is_cuda:
        raise LgcnError(f"{name} must be a CUDA tensor (got {t.device}); this package has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise LgcnError(f"{name} must be {dtype} (got {t.dtype})")
    return t


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Graph:
    """Device-resident CSR pair + normalisation + warp task lists for one edge list (K0).

    Built once per ``edge_index`` tensor and cached by callers; owns every array as a torch
    tensor and exposes the C struct the kernels take."""

    def __init__(self, edge_index: torch.Tensor, num_users: int, num_items: int):
        require_cuda(edge_index, "edge_index", torch.int64)
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise LgcnError(f"edge_index must be [2,E], got {tuple(edge_index.shape)}")
        ei = edge_index.contiguous()
        dev = ei.device
        n, e = num_users + num_items, ei.size(1)
        L = lib()
        sz = CGraphSizes()
        check(L.lgcn_graph_sizes_query(n, e, byref(sz)))
        i32 = dict(dtype=torch.int32, device=dev)
        self.device = dev
        self.num_users, self.num_items, self.num_nodes, self.num_edges = num_users, num_items, n, e
        e1 = max(e, 1)
        self.in_ptr = torch.empty(n + 1, **i32)
        self.out_ptr = torch.empty(n + 1, **i32)
        self.in_nbr, self.in_trip = torch.empty(e1, **i32), torch.empty(e1, **i32)
        self.out_nbr, self.out_trip = torch.empty(e1, **i32), torch.empty(e1, **i32)
        self.dis = torch.empty(n, dtype=torch.float32, device=dev)
        self.active = torch.empty(n, dtype=torch.uint8, device=dev)
        self.in_tasks = torch.empty(sz.task_bytes // 4, **i32)
        self.out_tasks = torch.empty(sz.task_bytes // 4, **i32)
        self.partials = torch.empty(sz.partial_bytes // 4, dtype=torch.float32, device=dev)
        self.slot_counters = torch.empty(sz.counter_bytes // 4, **i32)
        self.active_list = torch.empty(n, **i32)
        ws = torch.empty(sz.workspace_bytes, dtype=torch.uint8, device=dev)
        c = CGraph()
        c.in_ptr, c.in_nbr, c.in_trip = self.in_ptr.data_ptr(), self.in_nbr.data_ptr(), self.in_trip.data_ptr()
        c.out_ptr, c.out_nbr, c.out_trip = self.out_ptr.data_ptr(), self.out_nbr.data_ptr(), self.out_trip.data_ptr()
        c.dis, c.active = self.dis.data_ptr(), self.active.data_ptr()
        c.in_tasks, c.out_tasks = self.in_tasks.data_ptr(), self.out_tasks.data_ptr()
        c.partials, c.slot_counters = self.partials.data_ptr(), self.slot_counters.data_ptr()
        c.active_list = self.active_list.data_ptr()
        check(L.lgcn_graph_build(ei.data_ptr(), e, n, num_users, byref(c), ws.data_ptr(), ws.numel(),
                                 stream_ptr(dev)))
        del ws
        # shrink the over-allocated lists to what the build actually produced
        self.in_tasks = self.in_tasks[: max(c.n_in_tasks, 1) * 8].clone()
        self.out_tasks = self.out_tasks[: max(c.n_out_tasks, 1) * 8].clone()
        nslots = max(c.n_in_slots, c.n_out_slots, 1)
        self.partials = torch.empty(nslots * PARTIAL_STRIDE, dtype=torch.float32, device=dev)
        self.slot_counters = torch.zeros(nslots, **i32)
        self.active_list = self.active_list[: max(c.num_active, 1)].clone()
        c.active_list = self.active_list.data_ptr()
        self.sched = torch.zeros(2, **i32)
        c.sched = self.sched.data_ptr()
        c.in_tasks, c.out_tasks = self.in_tasks.data_ptr(), self.out_tasks.data_ptr()
        c.partials, c.slot_counters = self.partials.data_ptr(), self.slot_counters.data_ptr()
        self.c = c
        self.num_triplets = int(c.num_triplets)
        self.num_active = int(c.num_active)

    @property
    def ref(self):
        return byref(self.c)

    # integer views used by the bit-exact parity tests
    def in_degree(self) -> torch.Tensor:
        return (self.in_ptr[1:] - self.in_ptr[:-1]).to(torch.int64)

    def out_degree(self) -> torch.Tensor:
        return (self.out_ptr[1:] - self.out_ptr[:-1]).to(torch.int64)


class GraphView:
    """One graph of a BatchedGraphs build: the same duck type the step functions take from ``Graph``
    (``ref``, ``c``, the counts); its arrays are slices of the build's arena (``array(name)`` gives a
    tensor view for the parity tests)."""

    def __init__(self, owner: "BatchedGraphs", c: CGraph, num_users: int, num_items: int):
        self.owner, self.c = owner, c
        self.device = owner.device
        self.num_users, self.num_items, self.num_nodes = num_users, num_items, num_users + num_items
        self.num_edges = int(c.num_edges)
        self.num_triplets = int(c.num_triplets)
        self.num_active = int(c.num_active)

    @property
    def ref(self):
        return byref(self.c)

    def array(self, name: str) -> torch.Tensor:
        c, n, e = self.c, self.num_nodes, self.num_edges
        spec = {"in_ptr": (n + 1, torch.int32), "out_ptr": (n + 1, torch.int32),
                "in_nbr": (e, torch.int32), "in_trip": (e, torch.int32),
                "out_nbr": (e, torch.int32), "out_trip": (e, torch.int32),
                "dis": (n, torch.float32), "active": (n, torch.uint8),
                "in_tasks": (c.n_in_tasks * 8, torch.int32), "out_tasks": (c.n_out_tasks * 8, torch.int32),
                "active_list": (c.num_active, torch.int32)}[name]
        off = getattr(c, name) - self.owner.arena.data_ptr()
        nbytes = spec[0] * torch.empty((), dtype=spec[1]).element_size()
        return self.owner.arena[off: off + nbytes].view(spec[1])


class BatchedGraphs:
    """K0b: the graphs of B edge lists built by ONE ``lgcn_graph_build_batched`` call.

    ``edges``: device int64, the lists back to back (list b = its [2,E_b] tensor flattened at offset
    ``2*edge_off[b]``); ``edge_off``: B+1 ascending ints starting at 0.  ``arena`` / ``workspace``
    (uint8 CUDA tensors) are reused when large enough -- pass the previous build's to avoid
    re-allocating every epoch.  The graphs share scratch: use them one after another on one stream."""

    def __init__(self, edges: torch.Tensor, edge_off, num_users: int, num_items: int,
                 arena: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
        require_cuda(edges, "edges", torch.int64)
        dev = edges.device
        self.device = dev
        off = (c_int64 * len(edge_off))(*[int(x) for x in edge_off])
        b = len(edge_off) - 1
        n = num_users + num_items
        if edges.numel() < 2 * int(edge_off[-1]):
            raise LgcnError(f"edges holds {edges.numel()} ids, edge_off needs {2 * int(edge_off[-1])}")
        L = lib()
        sz = CBatchedSizes()
        check(L.lgcn_graph_batched_sizes(n, b, off, byref(sz)))
        if arena is None or arena.numel() < sz.arena_bytes:
            arena = torch.empty(int(sz.arena_bytes * 1.25) + 256, dtype=torch.uint8, device=dev)
        if workspace is None or workspace.numel() < sz.workspace_bytes:
            workspace = torch.empty(int(sz.workspace_bytes * 1.25) + 256, dtype=torch.uint8, device=dev)
        self.arena, self.workspace, self.edges = arena, workspace, edges
        cg = (CGraph * b)()
        check(L.lgcn_graph_build_batched(edges.data_ptr(), off, b, n, num_users, cg, arena.data_ptr(), arena.numel(),
                                         workspace.data_ptr(), workspace.numel(), stream_ptr(dev)))
        self._cg = cg
        self.graphs = [GraphView(self, cg[i], num_users, num_items) for i in range(b)]

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, i) -> GraphView:
        return self.graphs[i]


class StepBuffers:
    """Scratch for one forward/backward pass (reused across steps; sized for the largest P seen)."""

    def __init__(self, num_nodes: int, num_items: int, num_layers: int, device):
        f32 = dict(dtype=torch.float32, device=device)
        self.num_nodes, self.num_items, self.device = num_nodes, num_items, device
        self.final_emb = torch.empty(num_nodes, DIM, **f32)
        self.rnorm = torch.empty(num_nodes, **f32)
        self.grad_final = torch.empty(num_nodes, DIM, **f32)
        self.grad_e0 = torch.empty(num_nodes, DIM, **f32)
        self.work = torch.empty(max(num_layers - 1, 2) * num_nodes * DIM, **f32)
        self.neg_count = torch.zeros(num_items, dtype=torch.int32, device=device)
        self.accum = torch.zeros(4, dtype=torch.float64, device=device)
        self.trip_scratch = torch.empty(2, **f32)
        self.generation = 0
        # sparse steps: per-item stamp / list of the distinct inactive negatives of the current step
        self.neg_flag = torch.zeros(num_items, dtype=torch.int32, device=device)
        self.neg_list = torch.zeros(num_items, dtype=torch.int32, device=device)
        self.neg_list_count = torch.zeros(1, dtype=torch.int32, device=device)
        self.act_stamp = torch.zeros(2 * num_nodes + 4 * num_items, dtype=torch.int32, device=device)   # lgcn_train_steps_sparse
        self.steps_ws = torch.empty(0, dtype=torch.uint8, device=device)
        self.grad_final.zero_()            # the sparse step keeps dL/dfinal all-zero between steps
        self.c = CStepBuffers()
        self._fill()

    def _fill(self):
        c = self.c
        c.final_emb, c.rnorm = self.final_emb.data_ptr(), self.rnorm.data_ptr()
        c.grad_final, c.grad_e0 = self.grad_final.data_ptr(), self.grad_e0.data_ptr()
        c.work, c.work_bytes = self.work.data_ptr(), self.work.numel() * 4
        c.neg_count, c.accum = self.neg_count.data_ptr(), self.accum.data_ptr()
        c.trip_scratch = self.trip_scratch.data_ptr()
        c.neg_flag, c.neg_list = self.neg_flag.data_ptr(), self.neg_list.data_ptr()
        c.neg_list_count = self.neg_list_count.data_ptr()
        c.act_stamp = self.act_stamp.data_ptr()

    def ensure_triplets(self, p: int):
        if self.trip_scratch.numel() < 2 * p:
            self.trip_scratch = torch.empty(2 * p, dtype=torch.float32, device=self.device)
            self._fill()
            self.generation += 1          # captured CUDA graphs hold the old address

    @property
    def ref(self):
        return byref(self.c)
