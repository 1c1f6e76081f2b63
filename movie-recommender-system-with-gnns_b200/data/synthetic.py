"""Seeded synthetic MovieLens-shaped bipartite graphs.

Produces exactly what the reference's data pipeline hands to the hot path
(/root/reference/data/dataset_handler.py:111-141,160-185): users are ids ``0..U-1``, movies
``U..U+I-1``; ``edge_index`` is the ``to_undirected`` list sorted by (row, col) -- all
user->movie edges first, then all movie->user edges -- and the 90/5/5 split is taken over
DIRECTED edge positions (so the train graph is asymmetric), indices sorted ascending.

Generated with the torch CPU generator by default so that the CUDA path and the CPU oracle see
bit-identical inputs (every fixture and parity test uses that).  ``device=`` runs the same recipe with
that device's generator instead -- a DIFFERENT but equally shaped seeded graph, in seconds instead of
~90 s for the 10x shape; used only where the checker runs on the same in-process edge list
(bench.py --workload c5).  This is bench/test plumbing, not a kernel.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Tuple

import torch

# name -> (num_users, num_items, interactions, layers)   (BASELINE.json configs)
SHAPES = {
    "tiny": (200, 300, 6_000, 3),
    "ml100k": (943, 1_682, 100_000, 3),            # C1
    "ml1m": (6_040, 3_706, 575_000, 3),            # mid-size parity case
    "ml25m": (162_541, 59_047, 12_500_000, 3),     # C2 / C3 / C4
    "ml25m_x10": (1_600_000, 600_000, 125_000_000, 4),  # C5
}


@dataclass
class SyntheticGraph:
    num_users: int
    num_items: int
    edge_index: torch.Tensor      # [2, 2E'] int64, to_undirected order
    train_idx: torch.Tensor       # sorted positions into edge_index
    val_idx: torch.Tensor
    test_idx: torch.Tensor

    @property
    def num_nodes(self) -> int:
        return self.num_users + self.num_items

    def edges(self, split: str) -> torch.Tensor:
        idx = {"train": self.train_idx, "val": self.val_idx, "test": self.test_idx}[split]
        return self.edge_index[:, idx].contiguous()


def _zipf_mandelbrot_cdf(n: int, s: float, q: float, device=None) -> torch.Tensor:
    w = (torch.arange(1, n + 1, dtype=torch.float64, device=device) + q).pow(-s)
    return torch.cumsum(w / w.sum(), 0)


def _sample_pairs(count: int, cdf_u: torch.Tensor, cdf_i: torch.Tensor, perm_u: torch.Tensor,
                  perm_i: torch.Tensor, gen: torch.Generator) -> torch.Tensor:
    ru = torch.rand(count, generator=gen, dtype=torch.float64, device=cdf_u.device)
    ri = torch.rand(count, generator=gen, dtype=torch.float64, device=cdf_u.device)
    u = perm_u[torch.searchsorted(cdf_u, ru).clamp_(max=cdf_u.numel() - 1)]
    i = perm_i[torch.searchsorted(cdf_i, ri).clamp_(max=cdf_i.numel() - 1)]
    return u * cdf_i.numel() + i


def make_interactions(num_users: int, num_items: int, count: int, seed: int = 0, device=None) -> torch.Tensor:
    """``count`` unique (user, item) pairs, power-law on both sides (Zipf-Mandelbrot; item
    exponent 1.0, user exponent 0.8, heads flattened so no node exceeds the opposite side's
    size), every user and every item present at least once.  Returns keys ``u * I + i``
    sorted ascending."""
    assert count >= max(num_users, num_items) and count <= num_users * num_items
    dev = torch.device("cpu") if device is None else torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    cdf_u = _zipf_mandelbrot_cdf(num_users, 0.8, 50.0, dev)
    cdf_i = _zipf_mandelbrot_cdf(num_items, 1.0, 25.0, dev)
    perm_u = torch.randperm(num_users, generator=gen, device=dev)     # popularity rank -> id
    perm_i = torch.randperm(num_items, generator=gen, device=dev)
    # coverage: one pair per user and one per item (dataset_handler.py:111-112 counts only
    # entities that appear)
    cov_u = torch.arange(num_users, device=dev) * num_items + torch.randint(0, num_items, (num_users,), generator=gen, device=dev)
    cov_i = torch.randint(0, num_users, (num_items,), generator=gen, device=dev) * num_items + torch.arange(num_items, device=dev)
    keys = torch.unique(torch.cat([cov_u, cov_i]))
    while keys.numel() < count:
        need = count - keys.numel()
        extra = _sample_pairs(int(need * 1.3) + 1024, cdf_u, cdf_i, perm_u, perm_i, gen)
        extra = torch.unique(extra)
        extra = extra[~torch.isin(extra, keys)]
        if extra.numel() > need:                            # keep a seeded random subset
            extra = extra[torch.randperm(extra.numel(), generator=gen, device=dev)[:need]]
        keys = torch.unique(torch.cat([keys, extra]))
    return keys


def undirected_edge_index(keys: torch.Tensor, num_users: int, num_items: int) -> torch.Tensor:
    """The result ``to_undirected`` gives for a bipartite list with users first
    (dataset_handler.py:141): sorted by (row, col) and duplicate-free by construction."""
    u = keys // num_items
    m = keys % num_items + num_users
    # user->movie half: keys are already sorted by (u, m)
    back = torch.sort(m * (num_users + num_items) + u)[0]
    n = num_users + num_items
    row = torch.cat([u, back // n])
    col = torch.cat([m, back % n])
    return torch.stack([row, col])


def make_graph(shape: str = "ml100k", seed: int = 0, train_size: float = 0.9, device=None) -> SyntheticGraph:
    num_users, num_items, count, _ = SHAPES[shape]
    return make_graph_custom(num_users, num_items, count, seed, train_size, device)


def make_graph_custom(num_users: int, num_items: int, count: int, seed: int = 0,
                      train_size: float = 0.9, device=None) -> SyntheticGraph:
    """device=None: the CPU generator (the graph every fixture / parity test refers to); a CUDA device: the same
    recipe drawn from that device's generator (another graph of the same shape), tensors returned on the CPU."""
    dev = torch.device("cpu") if device is None else torch.device(device)
    keys = make_interactions(num_users, num_items, count, seed, dev)
    ei = undirected_edge_index(keys, num_users, num_items)
    del keys
    e = ei.shape[1]
    gen = torch.Generator(device=dev).manual_seed(seed + 1_000_003)
    perm = torch.randperm(e, generator=gen, device=dev)
    n_train = int(round(e * train_size))
    n_val = (e - n_train) // 2
    train_idx = torch.sort(perm[:n_train])[0]
    val_idx = torch.sort(perm[n_train:n_train + n_val])[0]
    test_idx = torch.sort(perm[n_train + n_val:])[0]
    return SyntheticGraph(num_users, num_items, ei.cpu(), train_idx.cpu(), val_idx.cpu(), test_idx.cpu())


def init_embeddings(num_users: int, num_items: int, dim: int = 64, seed: int = 0
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """N(0, 0.01^2) fp32 tables (models/light_gcn.py:25-26), CPU generator."""
    gen = torch.Generator().manual_seed(seed)
    u = torch.empty(num_users, dim).normal_(0.0, 0.01, generator=gen)
    i = torch.empty(num_items, dim).normal_(0.0, 0.01, generator=gen)
    return u, i


def hash_partition(num_nodes: int, num_parts: int) -> torch.Tensor:
    """Deterministic stand-in for METIS (decouples kernel timing from partitioner quality)."""
    x = torch.arange(num_nodes, dtype=torch.int64)
    x = (x * 2654435761) % (2 ** 32)
    x = x ^ (x >> 15)
    return x % num_parts


def shared_train_edges(shape: str, local_rank: int, barrier, device=None):
    """(num_users, num_items, train [2,E] int64 CPU tensor, SyntheticGraph or None).  Under torchrun only local rank 0
    generates the graph (the 10x graph takes ~90 s with the CPU generator and ~15 GB of host memory while it is being
    built; ``device``: generate there instead, see make_graph_custom); the other ranks read the train edges from
    /dev/shm after ``barrier()``."""
    import numpy as np
    nu, ni = SHAPES[shape][0], SHAPES[shape][1]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        g = make_graph(shape, seed=0, device=device)
        return g.num_users, g.num_items, g.edges("train"), g
    path = f"/dev/shm/lgcn_b200_{shape}_seed0_train_{os.environ.get('MASTER_PORT', '0')}.npy"
    if local_rank == 0:
        g = make_graph(shape, seed=0, device=device)
        np.save(path + ".tmp.npy", g.edges("train").numpy().astype(np.int32))
        os.replace(path + ".tmp.npy", path)
        del g
    barrier()
    tr = torch.from_numpy(np.load(path)).to(torch.int64)
    barrier()
    if local_rank == 0:
        os.remove(path)
    return nu, ni, tr, None
