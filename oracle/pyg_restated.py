"""ORACLE (test infrastructure, not product code) -- restatement of the third-party
functions the reference calls on its hot path.  **Parity unpinned** for this file: the
packages (torch-geometric==2.4.0, /root/reference/environment.yml:24; pytorch-sparse,
/root/reference/README.md:31) are not present in /root/reference nor installable offline, so
the algorithms below follow the published PyG 2.4.0 sources:

  torch_geometric/nn/conv/lg_conv.py      LGConv.forward / message
  torch_geometric/nn/conv/gcn_conv.py     gcn_norm (Tensor edge_index branch)
  torch_geometric/utils/scatter.py        scatter(reduce='sum')
  torch_geometric/utils/undirected.py     to_undirected
  torch_geometric/utils/coalesce.py       coalesce
  torch_geometric/utils/sort_edge_index.py, utils/sparse.py (index2ptr / ptr2index)
  torch_geometric/loader/cluster.py       ClusterData._metis/_partition/__getitem__
  torch_sparse/csrc/cpu/metis_cpu.cpp     partition -> METIS_PartGraphKway (idx_t = int64)

Reference call sites: models/light_gcn.py:4,24,33; data/dataset_handler.py:7-9,141,273,278,285.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterator, List, Optional

import torch

# --------------------------------------------------------------------------------------
# scatter / gcn_norm / LGConv
# --------------------------------------------------------------------------------------


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """utils/scatter.py: ``src.new_zeros(size).scatter_add_(0, broadcast(index), src)``."""
    size = list(src.shape)
    size[0] = dim_size
    if src.dim() > 1:
        index = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return src.new_zeros(size).scatter_add_(0, index, src)


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype: torch.dtype):
    """gcn_conv.py::gcn_norm with edge_weight=None, add_self_loops=False,
    flow='source_to_target'.  Returns (deg, deg_inv_sqrt, edge_weight)."""
    row, col = edge_index[0], edge_index[1]
    w = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    deg = scatter_sum(w, col, num_nodes)                       # IN-degree by target
    deg_inv_sqrt = deg.pow(-0.5)
    deg_inv_sqrt = deg_inv_sqrt.masked_fill(deg_inv_sqrt == float("inf"), 0)
    edge_weight = deg_inv_sqrt[row] * w * deg_inv_sqrt[col]
    return deg, deg_inv_sqrt, edge_weight


def lgconv(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """One LGConv layer: out[col[e]] += dis[row[e]] * dis[col[e]] * x[row[e]]  (uncached
    normalisation, gather -> mul -> scatter_add_, exactly the op sequence PyG issues)."""
    n = x.size(0)
    _, _, w = gcn_norm(edge_index, n, x.dtype)
    x_j = x.index_select(0, edge_index[0])
    msg = w.view(-1, 1) * x_j
    return scatter_sum(msg, edge_index[1], n)


class LGConv(torch.nn.Module):
    """Parameter-free, buffer-free stand-in with the call signature the reference uses
    (``conv(x=emb, edge_index=edge_index)``, models/light_gcn.py:33)."""

    def __init__(self, normalize: bool = True, **kwargs):
        super().__init__()
        assert normalize

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        return lgconv(x, edge_index)


# --------------------------------------------------------------------------------------
# coalesce / to_undirected / sort_edge_index / ptr helpers
# --------------------------------------------------------------------------------------


def coalesce(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    """utils/coalesce.py: key = row*N + col, stable ascending sort, drop keys equal to the
    predecessor."""
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    key = edge_index[0] * num_nodes + edge_index[1]
    key_sorted, perm = torch.sort(key, stable=True)
    edge_index = edge_index[:, perm]
    mask = torch.ones_like(key_sorted, dtype=torch.bool)
    mask[1:] = key_sorted[1:] > key_sorted[:-1]
    return edge_index[:, mask]


def to_undirected(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    """utils/undirected.py: cat (r,c)+(c,r) then coalesce."""
    row = torch.cat([edge_index[0], edge_index[1]])
    col = torch.cat([edge_index[1], edge_index[0]])
    return coalesce(torch.stack([row, col]), num_nodes)


def sort_edge_index(edge_index: torch.Tensor, num_nodes: int):
    """sort_by_row=True: perm = argsort(row*N + col) (stable).  Returns (edge_index, perm)."""
    key = edge_index[0] * num_nodes + edge_index[1]
    _, perm = torch.sort(key, stable=True)
    return edge_index[:, perm], perm


def index2ptr(index: torch.Tensor, size: int) -> torch.Tensor:
    return torch._convert_indices_from_coo_to_csr(index, size)


def ptr2index(ptr: torch.Tensor) -> torch.Tensor:
    ind = torch.arange(ptr.numel() - 1, dtype=ptr.dtype, device=ptr.device)
    return ind.repeat_interleave(ptr.diff())


# --------------------------------------------------------------------------------------
# METIS (torch_sparse.partition) -- realistic partition vectors for fixtures
# --------------------------------------------------------------------------------------

_HERE = os.path.dirname(os.path.abspath(__file__))
_METIS_SO = os.path.join(_HERE, "_ref", "libmetis_shim.so")


def metis_partition(indptr: torch.Tensor, index: torch.Tensor, num_parts: int) -> torch.Tensor:
    """torch.ops.torch_sparse.partition(indptr, index, None, num_parts, recursive=False):
    METIS_PartGraphKway, ncon=1, no weights, default options, 64-bit idx_t.  Needs
    oracle/_ref/libmetis_shim.so (oracle/Makefile)."""
    lib = ctypes.CDLL(_METIS_SO)
    lib.oracle_metis_kway.restype = ctypes.c_int
    lib.oracle_metis_kway.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_int64, ctypes.c_void_p]
    indptr = indptr.to(torch.int64).contiguous().clone()
    index = index.to(torch.int64).contiguous().clone()
    n = indptr.numel() - 1
    part = torch.empty(n, dtype=torch.int64)
    rc = lib.oracle_metis_kway(n, indptr.data_ptr(), index.data_ptr(), num_parts, part.data_ptr())
    if rc != 1:  # METIS_OK
        raise RuntimeError(f"METIS_PartGraphKway failed rc={rc}")
    return part


# --------------------------------------------------------------------------------------
# Data / ClusterData / DataLoader (loader/cluster.py, PyG 2.4.0)
# --------------------------------------------------------------------------------------


class Data:
    """Minimal torch_geometric.data.Data: attribute bag with ``.to(device)``."""

    def __init__(self, edge_index: Optional[torch.Tensor] = None, num_nodes: Optional[int] = None,
                 **kwargs):
        self.edge_index = edge_index
        self.num_nodes = num_nodes
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class ClusterData:
    """loader/cluster.py::ClusterData with defaults recursive=False, save_dir=None,
    keep_inter_cluster_edges=False, sparse_format='csr'.  ``cluster`` (the METIS output) may
    be supplied; bit-exactness downstream is defined GIVEN that vector."""

    def __init__(self, data: Data, num_parts: int, cluster: Optional[torch.Tensor] = None):
        self.num_parts = num_parts
        n = data.num_nodes
        ei = data.edge_index.cpu()
        if cluster is None:
            # _metis: row, index = sort_edge_index(edge_index, N); indptr = index2ptr(row, N)
            sorted_ei, _ = sort_edge_index(ei, n)
            indptr = index2ptr(sorted_ei[0], n)
            cluster = metis_partition(indptr, sorted_ei[1], num_parts)
        self.cluster = cluster.to(torch.int64).cpu()
        # _partition
        cluster_sorted, node_perm = torch.sort(self.cluster, stable=True)
        self.partptr = index2ptr(cluster_sorted, num_parts)
        self.node_perm = node_perm
        inv = torch.empty(n, dtype=torch.int64)
        inv[node_perm] = torch.arange(n, dtype=torch.int64)
        ei2 = inv[ei]
        ei2, self.edge_perm = sort_edge_index(ei2, n)
        self.indptr = index2ptr(ei2[0], n)
        self.index = ei2[1]
        # _permute_data: node-level attrs are reindexed by node_perm
        n_id = getattr(data, "n_id", None)
        self.n_id = (n_id.cpu()[node_perm] if n_id is not None else node_perm.clone())
        self._device = data.edge_index.device

    def __len__(self) -> int:
        return self.partptr.numel() - 1

    def __getitem__(self, p: int) -> Data:
        if p < 0 or p + 1 >= self.partptr.numel():
            raise IndexError(p)
        ns, ne = int(self.partptr[p]), int(self.partptr[p + 1])
        es, ee = int(self.indptr[ns]), int(self.indptr[ne])
        row = ptr2index(self.indptr[ns:ne + 1] - es)
        col = self.index[es:ee]
        mask = (col >= ns) & (col < ne)
        row, col = row[mask], col[mask] - ns
        out = Data(edge_index=torch.stack([row, col]).to(self._device), num_nodes=ne - ns)
        out.n_id = self.n_id[ns:ne].to(self._device)
        return out

    def __iter__(self) -> Iterator[Data]:
        p = 0
        while True:
            try:
                yield self[p]
            except IndexError:
                return
            p += 1


class DataLoader:
    """torch_geometric.loader.DataLoader(list_of_Data, batch_size=1, shuffle=True): a
    single-graph Batch keeps edge_index unchanged; order from torch.randperm (RandomSampler on
    the torch global generator)."""

    def __init__(self, dataset: List[Data], batch_size: int = 1, shuffle: bool = False):
        assert batch_size == 1
        self.dataset = list(dataset)
        self.shuffle = shuffle

    def __len__(self) -> int:
        return len(self.dataset)

    def __iter__(self) -> Iterator[Data]:
        n = len(self.dataset)
        if self.shuffle:
            # torch.utils.data.RandomSampler(generator=None): a seed is drawn from the global
            # generator, a private generator is seeded with it, then randperm.
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = range(n)
        for i in order:
            yield self.dataset[i]
