"""Triplet construction with the reference's signatures (/root/reference/utils/helpers.py:64-102).

In the fused training path these index vectors are never materialised (the kernels walk the
cached CSR instead); the functions exist for callers that compose the reference's pieces by hand.
The dead helpers of the reference (cantor_hash_pair, get_user_items, is_in_feasible -- never
called) are not reproduced.
"""
from __future__ import annotations

from typing import Tuple

import torch


def sample_negative(pos_idx: torch.Tensor, num_items: int, device: torch.device) -> torch.Tensor:
    """One uniform item id per positive, no rejection of true positives (utils/helpers.py:79-80).
    RNG = torch's generator of ``device`` exactly as in the reference."""
    return torch.randint(0, num_items, (pos_idx.shape[0],), device=device)


def get_triplets_indices(edge_index: torch.Tensor, num_users: int, num_items: int,
                         device: torch.device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(user, positive item, negative item) per directed user->movie edge, in edge order
    (utils/helpers.py:98-100)."""
    src, dst = edge_index[0], edge_index[1]
    users = src[src < num_users]
    pos_items = dst[dst >= num_users] - num_users
    return users, pos_items, sample_negative(pos_items, num_items, device)
