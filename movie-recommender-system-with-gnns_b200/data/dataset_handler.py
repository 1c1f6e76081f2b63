"""Graph container + Cluster-GCN batching with the reference's handler API
(/root/reference/data/dataset_handler.py:66-298).

``MovieLensDataHandler(ratings_path, movies_path)`` keeps the reference's attributes
(``edge_index, user_id_map, movie_id_map, id_user_map, id_movie_map, movies, num_users,
num_movies``) and methods (``get_datasets``, ``get_data_training``, ``get_num_users_items``).
``GraphDataHandler`` is the same object built from an in-memory edge list (synthetic
MovieLens-shaped graphs; no CSV).  The dataset download (dataset_handler.py:26-64) is out of scope.

Cluster-GCN batching (``get_data_training``) = METIS on the host (the call torch_sparse makes for
PyG's ClusterData) + ONE GPU extraction pass for all parts (``lgcn_cluster_extract``) instead of a
Python loop of boolean masks; the batches are views into one device buffer and carry GLOBAL ids with
``num_nodes = N`` exactly as dataset_handler.py:277-282 leaves them.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

from .._lib import LgcnError, check, lib, require_cuda, stream_ptr


class Data:
    """What the training loop needs from a PyG ``Data``/``Batch``: ``.edge_index``, ``.num_nodes``,
    ``.n_id`` and ``.to(device)`` (utils/train_test.py:87,98,120)."""

    def __init__(self, edge_index: Optional[torch.Tensor] = None, num_nodes: Optional[int] = None, **kw):
        self.edge_index = edge_index
        self.num_nodes = num_nodes
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        device = torch.device(device)
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v) and v.device != device:
                setattr(self, k, v.to(device))
        return self


class ClusterLoader:
    """``DataLoader(list_of_Data, batch_size=1, shuffle=True)`` (dataset_handler.py:285): yields the
    SAME Data objects every epoch (so per-batch CSR caches hit), in an order drawn the way torch's
    RandomSampler does from the global generator."""

    def __init__(self, dataset: List[Data], shuffle: bool = True):
        self.dataset = list(dataset)
        self.shuffle = shuffle

    def __len__(self) -> int:
        return len(self.dataset)

    def __iter__(self) -> Iterator[Data]:
        n = len(self.dataset)
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = range(n)
        for i in order:
            yield self.dataset[i]


def to_undirected(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    """PyG ``to_undirected`` (dataset_handler.py:141): both directions, sorted by (row, col),
    duplicates dropped -- ``lgcn_to_undirected``: key build, one 64-bit radix sort, head-flag scan and
    compaction on the device (no CPU path: the edges must be a CUDA tensor)."""
    require_cuda(edge_index, "edge_index", torch.int64)
    ei = edge_index.contiguous()
    dev = ei.device
    e = ei.size(1)
    if e == 0:
        return torch.empty(2, 0, dtype=torch.int64, device=dev)
    n = int(ei.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.empty(4 * e, dtype=torch.int64, device=dev)
    ws = torch.empty(lib().lgcn_to_undirected_workspace_bytes(e), dtype=torch.uint8, device=dev)
    count = ctypes.c_int64(0)
    check(lib().lgcn_to_undirected(ei.data_ptr(), e, n, out.data_ptr(), ctypes.addressof(count), ws.data_ptr(),
                                   ws.numel(), stream_ptr(dev)))
    return out[:2 * count.value].view(2, count.value)


def shuffle_split(n: int, train_size: Optional[float] = None, test_size: Optional[float] = None
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """``sklearn.model_selection.train_test_split(np.arange(n), train_size=..|test_size=.., shuffle=True)`` as the
    reference calls it (dataset_handler.py:167-168), restated: ``n_test = ceil(test_size*n)`` / ``n_train =
    floor(train_size*n)`` (the missing one is the complement), ONE permutation from numpy's global RandomState,
    test = its first n_test positions, train = the next n_train.  Consumes the same random stream as sklearn, so
    a seeded run splits identically (tests/test_host_logic_cpu.py compares with sklearn itself and the fixture)."""
    if (train_size is None) == (test_size is None):
        raise ValueError("give exactly one of train_size / test_size")
    if test_size is not None:
        n_test = int(np.ceil(test_size * n))
        n_train = n - n_test
    else:
        n_train = int(np.floor(train_size * n))
        n_test = n - n_train
    if n_train <= 0 or n_test <= 0:
        raise ValueError(f"With n_samples={n}, the resulting train/test set would be empty")
    perm = np.random.permutation(n)
    return perm[n_test:n_test + n_train], perm[:n_test]


def metis_partition(edge_index: torch.Tensor, num_nodes: int, num_parts: int) -> torch.Tensor:
    """ClusterData._metis: CSR of the (directed, possibly asymmetric) edge list sorted by
    (row, col), then METIS_PartGraphKway with torch_sparse's arguments.  Host call."""
    ei = edge_index.cpu()
    key = torch.sort(ei[0] * num_nodes + ei[1], stable=True)[0]
    row, col = key // num_nodes, (key % num_nodes).contiguous()
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(torch.bincount(row, minlength=num_nodes), 0)
    part = torch.empty(num_nodes, dtype=torch.int64)
    rc = lib().lgcn_partition_metis(num_nodes, indptr.data_ptr(), col.data_ptr(), num_parts, part.data_ptr())
    if rc != 0:
        raise LgcnError(f"METIS failed (rc={rc})")
    return part


def cluster_extract(edge_index: torch.Tensor, num_nodes: int, cluster: torch.Tensor, num_parts: int
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K4 on the device: returns (edges [2,kept] int64 GLOBAL ids grouped by part, part_ptr [P+1])."""
    require_cuda(edge_index, "edge_index", torch.int64)
    ei = edge_index.contiguous()
    dev = ei.device
    cl = cluster.to(device=dev, dtype=torch.int64).contiguous()
    if cl.numel() != num_nodes:
        raise LgcnError(f"cluster has {cl.numel()} entries for {num_nodes} nodes")
    e = ei.size(1)
    out = torch.empty(2, max(e, 1), dtype=torch.int64, device=dev)
    part_ptr = torch.empty(num_parts + 1, dtype=torch.int64, device=dev)
    ws = torch.empty(lib().lgcn_cluster_extract_workspace_bytes(num_nodes, e, num_parts), dtype=torch.uint8, device=dev)
    check(lib().lgcn_cluster_extract(ei.data_ptr(), e, num_nodes, cl.data_ptr(), num_parts, out.data_ptr(),
                                     part_ptr.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)))
    kept = int(part_ptr[-1])
    return out[:, :kept], part_ptr


class ClusterData:
    """PyG ``ClusterData(data, num_parts)`` followed by the reference's global-id remap
    (dataset_handler.py:273-282), all parts at once.  Iterating yields ``Data`` with GLOBAL ids,
    ``num_nodes = N`` and ``n_id = arange(N)``."""

    def __init__(self, data: Data, num_parts: int, cluster: Optional[torch.Tensor] = None,
                 partitioner: str = "metis", num_users: Optional[int] = None):
        """``cluster``: a ready partition vector; else ``partitioner`` computes one: "metis" = the host call PyG makes
        (the reference's partitions), "gpu" = balanced label propagation on the device (data/partition_gpu.py: other
        partitions, no host stage; ``num_users`` tells it the bipartite sides)."""
        n = data.num_nodes
        self.num_parts = num_parts
        if cluster is not None:
            self.cluster = cluster
        elif partitioner == "metis":
            self.cluster = metis_partition(data.edge_index, n, num_parts)
        elif partitioner == "gpu":
            from .partition_gpu import gpu_partition
            self.cluster = gpu_partition(data.edge_index, n, num_parts, num_users)
        else:
            raise ValueError(f"unknown partitioner {partitioner!r} (metis | gpu)")
        edges, part_ptr = cluster_extract(data.edge_index, n, self.cluster, num_parts)
        self.part_ptr = part_ptr.cpu()
        n_id = torch.arange(n, device=edges.device)
        self.parts: List[Data] = []
        for p in range(num_parts):
            b, e = int(self.part_ptr[p]), int(self.part_ptr[p + 1])
            self.parts.append(Data(edge_index=edges[:, b:e].contiguous(), num_nodes=n, n_id=n_id))

    def __len__(self) -> int:
        return self.num_parts

    def __iter__(self) -> Iterator[Data]:
        return iter(self.parts)

    def __getitem__(self, p: int) -> Data:
        return self.parts[p]


class GraphDataHandler:
    """The reference handler's post-CSV state, from an in-memory ``to_undirected`` edge list."""

    def __init__(self, edge_index: torch.Tensor, num_users: int, num_movies: int,
                 device: Optional[torch.device] = None):
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.edge_index = edge_index
        self.num_users, self.num_movies = num_users, num_movies
        self.user_id_map: Dict[int, int] = {}
        self.movie_id_map: Dict[int, int] = {}
        self._split: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None

    def set_split(self, train_idx, val_idx, test_idx) -> None:
        self._split = tuple(np.asarray(torch.as_tensor(x).cpu()) for x in (train_idx, val_idx, test_idx))

    def _indices(self, train_size: float):
        if self._split is None:
            # dataset_handler.py:167-172: two shuffle splits over DIRECTED edge positions (90 % train, the rest halved
            # into val / test), each index set sorted
            train, val_test = shuffle_split(self.edge_index.shape[1], train_size=train_size)
            vt_train, vt_test = shuffle_split(len(val_test), test_size=0.5)
            val, test = val_test[vt_train], val_test[vt_test]
            self._split = (np.sort(train), np.sort(val), np.sort(test))
        return self._split

    def get_datasets(self, train_size: float = 0.9) -> Tuple[Data, Data, Data]:
        tr, va, te = self._indices(train_size)
        n = self.num_users + self.num_movies
        out = []
        for idx in (tr, va, te):
            ei = self.edge_index[:, torch.as_tensor(idx)].contiguous()
            d = Data(edge_index=ei, num_nodes=n).to(self.device)
            d.n_id = torch.arange(n, device=self.device)
            out.append(d)
        return tuple(out)

    def get_data_training(self, num_train_clusters: int = 100, cluster: Optional[torch.Tensor] = None,
                          partitioner: str = "metis") -> Tuple[ClusterLoader, Data, Data]:
        train, val, test = self.get_datasets()
        cd = ClusterData(train, num_parts=num_train_clusters, cluster=cluster, partitioner=partitioner,
                         num_users=self.num_users)
        return ClusterLoader(cd.parts, shuffle=True), val, test

    def get_num_users_items(self) -> Tuple[int, int]:
        return self.num_users, self.num_movies


class MovieLensDataHandler(GraphDataHandler):
    """CSV front end (dataset_handler.py:98-141): ratings >= 4 are edges, ids mapped in order of
    first appearance (users 0..U-1, movies U..U+I-1), then ``to_undirected``.  Split indices are
    persisted / reloaded as ``data/indexes/{val,test}_indices.npy`` like the reference (:155-253)."""

    def __init__(self, ratings_path: str, movies_path: str, device: Optional[torch.device] = None):
        import pandas as pd
        if not os.path.exists(ratings_path) or not os.path.exists(movies_path):
            raise FileNotFoundError(f"{ratings_path} / {movies_path} not found (no download in this build)")
        ratings = pd.read_csv(ratings_path, usecols=['userId', 'movieId', 'rating'])
        ratings = ratings[ratings['rating'] >= 4]
        self.movies = pd.read_csv(movies_path, usecols=['movieId', 'title'])
        users, movies = ratings['userId'].unique(), ratings['movieId'].unique()
        nu = len(users)
        user_id_map = {int(k): i for i, k in enumerate(users)}
        movie_id_map = {int(k): i + nu for i, k in enumerate(movies)}
        u = ratings['userId'].map(user_id_map).values
        m = ratings['movieId'].map(movie_id_map).values
        dev = torch.device("cuda") if device is None else torch.device(device)
        ei = torch.from_numpy(np.vstack((u, m))).long().to(dev)
        super().__init__(to_undirected(ei).cpu(), nu, len(movies), dev)
        self.ratings_path, self.movies_path = ratings_path, movies_path
        self.user_id_map, self.movie_id_map = user_id_map, movie_id_map
        self.id_user_map = {i: k for k, i in user_id_map.items()}
        self.id_movie_map = {i: k for k, i in movie_id_map.items()}

    def _indices(self, train_size: float):
        if self._split is not None:
            return self._split
        path = "data/indexes"
        vf, tf = os.path.join(path, "val_indices.npy"), os.path.join(path, "test_indices.npy")
        e = self.edge_index.shape[1]
        if os.path.exists(path):
            va, te = np.sort(np.load(vf)), np.sort(np.load(tf))
            tr = np.setdiff1d(np.arange(e), np.concatenate((va, te)))
            self._split = (tr, va, te)
        else:
            tr, va, te = super()._indices(train_size)
            os.makedirs(path)
            np.save(vf, va)
            np.save(tf, te)
        return self._split

    def get_num_users_items(self) -> Tuple[int, int]:
        return len(self.user_id_map), len(self.movie_id_map)
