"""GPU graph partitioner for Cluster-GCN batching -- the alternative to the host METIS call
PyG's ``ClusterData`` makes (/root/reference/data/dataset_handler.py:273).

NOT a parity item: the partition vector differs from METIS', so batches, losses and weights differ too; what it trades
is set-up time (METIS: ~75 s on the host at ML-25M shape) against the share of edges that stay inside a part (the
edges Cluster-GCN trains on).  Everything downstream (``lgcn_cluster_extract``, the step kernels) is unchanged and
stays bit-exact GIVEN the vector.

Algorithm -- balanced label propagation, deterministic, integer only:
  start    two candidate labelings: a hash of the node id, and degree bands per side (the p-th band of the users and
           the p-th band of the items by descending degree share part p -- near-optimal for a rank-1 / Chung-Lu-like
           bipartite graph, where the hubs of both sides belong together);
  round    for the item side, then the user side: every node votes for the label most of its out-neighbours carry
           (``lgcn_label_vote``: one warp per row, shared-memory histogram, arg-max with ties to the smallest label);
           moves are accepted per target part in order of (gain desc, id asc) while the part has room under the
           round's capacity ``ceil(N/P * (1 + eps))`` -- repeated until nothing moves (the votes stay valid: they
           depend on the OTHER side's labels only);
  evict    parts above the capacity give up their least attached nodes (fewest out-neighbours in the own part) to the
           parts with room;
  schedule eps = 1, 1, 0.5, 0.25, 0.1, imb, imb: loose capacities first so that clusters can form, then tightened to the
           requested imbalance (METIS' default kway imbalance is 1.03);
  result   the refined labeling with more intra-part edges.

The orchestration below is device-agnostic torch (sorts, bincounts); the product entry point ``gpu_partition`` requires
CUDA tensors and uses the voting kernel; ``tests/partition_ref.py`` plugs a torch restatement of the vote into the same
orchestration to check the kernel bit-exactly.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence, Tuple

import torch

from .._lib import LgcnError, check, lib, require_cuda, stream_ptr

SCHEDULE = (1.0, 1.0, 0.5, 0.25, 0.1)
_ID_BITS, _GAIN_BITS = 22, 20          # sort key = part << 42 | (2^20 - 1 - gain) << 22 | local id


def csr_by_source(edge_index: torch.Tensor, num_nodes: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(ptr [N+1] int32, nbr [E] int32) of the directed edge list sorted by (row, col) -- the adjacency METIS gets."""
    key = torch.sort(edge_index[0] * num_nodes + edge_index[1])[0]
    row = key // num_nodes
    ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=edge_index.device)
    ptr[1:] = torch.cumsum(torch.bincount(row, minlength=num_nodes), 0)
    return ptr.to(torch.int32), (key % num_nodes).to(torch.int32)


def hash_labels(num_nodes: int, num_parts: int, device) -> torch.Tensor:
    x = torch.arange(num_nodes, dtype=torch.int64, device=device)
    x = (x * 2654435761) % (2 ** 32)
    x = x ^ (x >> 15)
    return x % num_parts


def band_labels(degree: torch.Tensor, sides: Sequence[Tuple[int, int]], num_parts: int) -> torch.Tensor:
    out = torch.empty_like(degree)
    for b, e in sides:
        order = torch.sort(degree[b:e], descending=True, stable=True)[1]
        rank = torch.empty_like(order)
        rank[order] = torch.arange(e - b, device=degree.device)
        out[b:e] = (rank * num_parts) // max(e - b, 1)
    return out


def cuda_vote(ptr: torch.Tensor, nbr: torch.Tensor, labels: torch.Tensor, b: int, e: int, num_parts: int):
    """(want, best, own) for rows [b,e) -- int64 tensors of length e-b -- from ``lgcn_label_vote``."""
    n = ptr.numel() - 1
    dev = ptr.device
    lab32 = labels.to(torch.int32)
    out = torch.zeros(3, n, dtype=torch.int32, device=dev)
    check(lib().lgcn_label_vote(ptr.data_ptr(), nbr.data_ptr(), lab32.data_ptr(), b, e, num_parts, out[0].data_ptr(),
                                out[1].data_ptr(), out[2].data_ptr(), stream_ptr(dev)))
    return out[0, b:e].long(), out[1, b:e].long(), out[2, b:e].long()


def _accept(labels: torch.Tensor, b: int, e: int, want: torch.Tensor, gain: torch.Tensor, num_parts: int, cap: int) -> int:
    dev = labels.device
    cur = labels[b:e]
    room = (cap - torch.bincount(labels, minlength=num_parts)).clamp(min=0)
    idx = torch.nonzero((want != cur) & (gain > 0)).squeeze(1)
    if idx.numel() == 0:
        return 0
    key = (want[idx] << (_ID_BITS + _GAIN_BITS)) | ((((1 << _GAIN_BITS) - 1) - gain[idx]) << _ID_BITS) | idx
    key = torch.sort(key)[0]
    w = key >> (_ID_BITS + _GAIN_BITS)
    ids = key & ((1 << _ID_BITS) - 1)
    start = torch.searchsorted(w, torch.arange(num_parts, device=dev))
    rank = torch.arange(key.numel(), device=dev) - start[w]
    ok = rank < room[w]
    labels[b + ids[ok]] = w[ok]
    return int(ok.sum())


def _evict(labels: torch.Tensor, own: torch.Tensor, num_parts: int, cap: int) -> int:
    dev, n = labels.device, labels.numel()
    size = torch.bincount(labels, minlength=num_parts)
    over = (size - cap).clamp(min=0)
    if int(over.sum()) == 0:
        return 0
    key = (labels << (_ID_BITS + _GAIN_BITS)) | (own << _ID_BITS) | torch.arange(n, device=dev)
    key = torch.sort(key)[0]
    w = key >> (_ID_BITS + _GAIN_BITS)
    ids = key & ((1 << _ID_BITS) - 1)
    start = torch.searchsorted(w, torch.arange(num_parts, device=dev))
    rank = torch.arange(n, device=dev) - start[w]
    out = ids[rank < over[w]]
    slots = torch.repeat_interleave(torch.arange(num_parts, device=dev), (cap - size).clamp(min=0))
    labels[out] = slots[: out.numel()]
    return int(out.numel())


def refine(ptr: torch.Tensor, nbr: torch.Tensor, labels: torch.Tensor, sides: Sequence[Tuple[int, int]], num_parts: int,
           imbalance: float, vote: Callable, schedule: Sequence[float] = SCHEDULE) -> torch.Tensor:
    n = labels.numel()
    labels = labels.clone()
    eps_list = list(schedule) + [imbalance, imbalance]
    cap_final = int(math.ceil(n / num_parts * (1.0 + imbalance)))
    for r, eps in enumerate(eps_list):
        cap = int(math.ceil(n / num_parts * (1.0 + max(eps, imbalance))))
        for b, e in sides:
            want, best, own = vote(ptr, nbr, labels, b, e, num_parts)
            gain = best - own
            for _ in range(8):
                if _accept(labels, b, e, want, gain, num_parts, cap) == 0:
                    break
        own_all = vote(ptr, nbr, labels, 0, n, num_parts)[2]
        _evict(labels, own_all, num_parts, cap_final if r >= len(eps_list) - 3 else cap)
    return labels


def partition(edge_index: torch.Tensor, num_nodes: int, num_parts: int, num_users: Optional[int] = None,
              imbalance: float = 0.03, vote: Callable = cuda_vote) -> Tuple[torch.Tensor, dict]:
    """(cluster [N] int64 on edge_index's device, stats).  ``num_users``: bipartite graph with users first (the two
    sides are updated alternately); None: one side, synchronous updates."""
    if num_nodes >= (1 << _ID_BITS) or num_parts > 4096:
        raise LgcnError(f"gpu partitioner handles up to {(1 << _ID_BITS) - 1} nodes and 4096 parts")
    dev = edge_index.device
    row, col = edge_index[0], edge_index[1]
    ptr, nbr = csr_by_source(edge_index, num_nodes)
    sides = [(num_users, num_nodes), (0, num_users)] if num_users else [(0, num_nodes)]
    degree = torch.bincount(row, minlength=num_nodes) + torch.bincount(col, minlength=num_nodes)
    if int(degree.max()) >= (1 << _GAIN_BITS):
        raise LgcnError("gpu partitioner: a node degree exceeds 2^20")
    best_labels, best_intra, tried = None, -1, {}
    for name, start in (("hash", hash_labels(num_nodes, num_parts, dev)),
                        ("degree_bands", band_labels(degree, sides if num_users else [(0, num_nodes)], num_parts))):
        lab = refine(ptr, nbr, start, sides, num_parts, imbalance, vote)
        intra = int((lab[row] == lab[col]).sum())
        tried[name] = intra
        if intra > best_intra:
            best_labels, best_intra = lab, intra
    size = torch.bincount(best_labels, minlength=num_parts)
    stats = {"intra_edges": best_intra, "intra_edge_share": best_intra / max(edge_index.shape[1], 1),
             "max_part": int(size.max()), "min_part": int(size.min()), "starts": tried}
    return best_labels, stats


def gpu_partition(edge_index: torch.Tensor, num_nodes: int, num_parts: int, num_users: Optional[int] = None,
                  imbalance: float = 0.03) -> torch.Tensor:
    """Product entry point: the partition vector for ``ClusterData(..., partitioner="gpu")``.  CUDA only."""
    require_cuda(edge_index, "edge_index", torch.int64)
    return partition(edge_index, num_nodes, num_parts, num_users, imbalance)[0]
