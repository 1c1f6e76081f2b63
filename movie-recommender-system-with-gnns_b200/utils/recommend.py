"""Serving / full-rank evaluation on the fused scoring kernel, with the reference's
``recommend_from_user`` / ``recommend_from_movie`` signatures and return values
(/root/reference/utils/recommend.py:12-113).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from .._lib import DIM, LgcnError, check, lib, require_cuda, stream_ptr


SCORE_AUTO, SCORE_FFMA, SCORE_TENSOR = 0, 1, 2


def score_topk(user_rows: torch.Tensor, item_rows: torch.Tensor, k: int, normalize: bool = True,
               excl_ptr: Optional[torch.Tensor] = None, excl_idx: Optional[torch.Tensor] = None,
               u_begin: int = 0, u_end: Optional[int] = None, algo: int = SCORE_AUTO,
               pack_items: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """top-k items per user by (score desc, id asc), score = <u, i> of (optionally L2-normalised)
    rows, items in the user's exclusion row removed.  ``excl_ptr`` [U+1] int64 / ``excl_idx`` int32
    sorted per row (CSR of train user->movie edges).  Returns (idx [n,k] int32, val [n,k])."""
    require_cuda(user_rows, "user_rows", torch.float32)
    require_cuda(item_rows, "item_rows", torch.float32)
    if user_rows.dim() != 2 or user_rows.size(1) != DIM or item_rows.dim() != 2 or item_rows.size(1) != DIM:
        raise LgcnError(f"score_topk expects [*,{DIM}] row tensors")
    ur, ir = user_rows.contiguous(), item_rows.contiguous()
    u_end = ur.size(0) if u_end is None else u_end
    n = u_end - u_begin
    dev = ur.device
    idx = torch.empty(n, k, dtype=torch.int32, device=dev)
    val = torch.empty(n, k, dtype=torch.float32, device=dev)
    ep = ex = None
    if excl_ptr is not None:
        ep = require_cuda(excl_ptr, "excl_ptr", torch.int64).contiguous()
        ex = require_cuda(excl_idx, "excl_idx", torch.int32).contiguous()
    ws = None
    if algo != SCORE_FFMA and k <= 32 and pack_items:       # tensor-core kernel: pre-packed item tiles fetched by TMA
        ws = torch.empty(lib().lgcn_score_topk_workspace_bytes(ir.size(0)), dtype=torch.uint8, device=dev)
    check(lib().lgcn_score_topk_ex(ur.data_ptr(), ir.data_ptr(), ir.size(0), u_begin, u_end, int(normalize),
                                   None if ep is None else ep.data_ptr(), None if ex is None else ex.data_ptr(),
                                   k, idx.data_ptr(), val.data_ptr(), algo, None if ws is None else ws.data_ptr(),
                                   0 if ws is None else ws.numel(), stream_ptr(dev)))
    return idx, val


def exclusion_csr(train_edge_index: torch.Tensor, num_users: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """CSR over users of their train items (the batched form of utils/recommend.py:141-142:
    ``edge_index[1, edge_index[0] == u] - num_users``), item ids sorted per user."""
    src, dst = train_edge_index[0], train_edge_index[1]
    m = src < num_users
    u, it = src[m], dst[m] - num_users
    key = torch.sort(u * (int(it.max()) + 1 if it.numel() else 1) + it)[0]
    stride = int(it.max()) + 1 if it.numel() else 1
    ptr = torch.zeros(num_users + 1, dtype=torch.int64, device=src.device)
    ptr[1:] = torch.cumsum(torch.bincount(key // stride, minlength=num_users), 0)
    return ptr, (key % stride).to(torch.int32)


def full_rank_eval(user_emb: torch.Tensor, item_emb: torch.Tensor, train_edge_index: torch.Tensor,
                   test_edge_index: torch.Tensor, num_users: int, k: int = 20, normalize: bool = True,
                   user_block: int = 32768, u_begin: int = 0, u_end: Optional[int] = None,
                   allreduce=None) -> Dict[str, float]:
    """BASELINE config C4: all-user x all-item scoring, train items masked, recall@k / NDCG@k against
    the held-out user->movie edges.  The score matrix is never materialised.

    Multi-GPU ("embarrassingly parallel over user ranges; items replicated"): every rank scores
    its own user range [u_begin, u_end) and ``allreduce`` (a callable that sums a float64 tensor over the ranks
    in place) combines the three partial sums -- see ``sharded_full_rank_eval``."""
    u_end = num_users if u_end is None else u_end
    dev = user_emb.device
    ptr, idx = exclusion_csr(train_edge_index, num_users)
    tops = []
    for b in range(u_begin, u_end, user_block):
        tops.append(score_topk(user_emb, item_emb, k, normalize, ptr, idx, b, min(u_end, b + user_block))[0])
    top = torch.cat(tops).to(torch.int64) if tops else torch.empty(0, k, dtype=torch.int64, device=dev)
    nloc = u_end - u_begin
    tptr, tidx = exclusion_csr(test_edge_index, num_users)
    cnt_all = (tptr[1:] - tptr[:-1])
    cnt = cnt_all[u_begin:u_end]
    # membership test: (user, item) keys of the held-out edges, sorted
    n_items = item_emb.size(0)
    owner = torch.repeat_interleave(torch.arange(num_users, device=dev), cnt_all)
    truth = torch.sort(owner * n_items + tidx.to(torch.int64))[0]
    keys = torch.arange(u_begin, u_end, device=dev).unsqueeze(1) * n_items + top.clamp(min=0)
    pos = torch.searchsorted(truth, keys.reshape(-1)).clamp(max=max(truth.numel() - 1, 0))
    hit = ((truth[pos] == keys.reshape(-1)) if truth.numel() else torch.zeros_like(pos, dtype=torch.bool))
    hit = (hit.reshape(nloc, k) & (top >= 0)).to(torch.float64)
    disc = 1.0 / torch.log2(torch.arange(2, k + 2, dtype=torch.float64, device=dev))
    has = cnt > 0
    idcg = torch.cumsum(disc, 0)[(cnt.clamp(max=k) - 1).clamp(min=0)]
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    if bool(has.any()):
        sums[0] = (hit.sum(1)[has] / cnt[has]).sum()
        sums[1] = ((hit * disc).sum(1)[has] / idcg[has]).sum()
        sums[2] = has.sum()
    if allreduce is not None:
        allreduce(sums)
    users = int(sums[2])
    return {"recall": float(sums[0]) / users if users else 0.0, "ndcg": float(sums[1]) / users if users else 0.0,
            "users": users}


def user_ranges(num_users: int, world: int, tile: int = 128) -> List[Tuple[int, int]]:
    """Equal user ranges for sharded scoring, cut at multiples of the scoring kernel's 128-user CTA tile (every user
    costs the same: one pass over all items)."""
    tiles = (num_users + tile - 1) // tile
    cuts = [min(num_users, ((tiles * r) // world) * tile) for r in range(world)] + [num_users]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def sharded_full_rank_eval(user_emb: torch.Tensor, item_emb: torch.Tensor, train_edge_index: torch.Tensor,
                           test_edge_index: torch.Tensor, num_users: int, k: int = 20, normalize: bool = True,
                           group=None) -> Dict[str, float]:
    """C4 over the GPUs of a ``torch.distributed`` job: rank r scores users ``user_ranges(...)[r]`` against the
    replicated item table; one all-reduce of three doubles yields the same recall / NDCG on every rank."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    lo, hi = user_ranges(num_users, world)[rank]

    def allreduce(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = full_rank_eval(user_emb, item_emb, train_edge_index, test_edge_index, num_users, k, normalize,
                         u_begin=lo, u_end=hi, allreduce=allreduce)
    out["user_range"] = (lo, hi)
    return out


def recommend_from_user(model: torch.nn.Module, user_id: int, data_handler: Any,
                        excluded_train_items: Optional[Union[List[int], torch.Tensor]] = None
                        ) -> Dict[str, Union[str, List[Dict[str, Union[str, float]]]]]:
    """Top-10 movies for ``user_id`` by cosine of LAYER-0 rows, skipping ``excluded_train_items``
    (utils/recommend.py:12-63).  Same return dicts, including ``{'error': 'Invalid user ID'}``."""
    user_index = data_handler.user_id_map.get(user_id)
    if user_index is None:
        return {'error': 'Invalid user ID'}
    with torch.no_grad():
        uw, iw = model.user_embedding.weight, model.item_embedding.weight
        dev = uw.device
        ptr = idx = None
        if excluded_train_items is not None:
            ex = torch.as_tensor(excluded_train_items, dtype=torch.int64, device=dev).reshape(-1)
            ex = torch.unique(ex)                                   # sorted
            ex = ex[(ex >= 0) & (ex < iw.size(0))]
            ptr = torch.zeros(uw.size(0) + 1, dtype=torch.int64, device=dev)
            ptr[user_index + 1:] = ex.numel()
            idx = ex.to(torch.int32)
        top_idx, top_val = score_topk(uw, iw, 10, True, ptr, idx, user_index, user_index + 1)
        top_idx, top_val = top_idx[0].tolist(), top_val[0].tolist()
    id_movie = getattr(data_handler, "id_movie_map", None)
    if id_movie is None:
        id_movie = {v: k for k, v in data_handler.movie_id_map.items()}
    n_users = len(data_handler.user_id_map)
    movies = data_handler.movies
    recommendations = []
    for i, s in zip(top_idx, top_val):
        if i < 0:
            break
        movie_id = id_movie[i + n_users]
        title = movies[movies['movieId'] == movie_id].iloc[0]['title']
        recommendations.append({'title': title, 'score': s})
    return {'recommendations': recommendations}


def recommend_from_movie(model: torch.nn.Module, movie_id: int, data_handler: Any,
                         excluded_train_users: Optional[Union[List[int], torch.Tensor]] = None
                         ) -> Dict[str, Union[str, List[Dict[str, Union[int, float]]]]]:
    """Top-10 users for ``movie_id`` (utils/recommend.py:65-113): the same kernel with the roles of
    the two tables swapped."""
    movie_index = data_handler.movie_id_map.get(movie_id)
    if movie_index is None:
        return {'error': 'Invalid movie ID'}
    n_users = len(data_handler.user_id_map)
    movie_index -= n_users
    with torch.no_grad():
        uw, iw = model.user_embedding.weight, model.item_embedding.weight
        dev = uw.device
        ptr = idx = None
        if excluded_train_users is not None:
            ex = torch.unique(torch.as_tensor(excluded_train_users, dtype=torch.int64, device=dev).reshape(-1))
            ex = ex[(ex >= 0) & (ex < uw.size(0))]
            ptr = torch.zeros(iw.size(0) + 1, dtype=torch.int64, device=dev)
            ptr[movie_index + 1:] = ex.numel()
            idx = ex.to(torch.int32)
        top_idx, top_val = score_topk(iw, uw, 10, True, ptr, idx, movie_index, movie_index + 1)
        top_idx, top_val = top_idx[0].tolist(), top_val[0].tolist()
    id_user = getattr(data_handler, "id_user_map", None) or {v: k for k, v in data_handler.user_id_map.items()}
    return {'top_users': [{'user_id': id_user[i], 'score': s} for i, s in zip(top_idx, top_val) if i >= 0]}
