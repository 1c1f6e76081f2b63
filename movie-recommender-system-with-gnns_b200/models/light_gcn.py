"""LightGCN module with the reference's API (/root/reference/models/light_gcn.py:13-64) on top of
the sm_100a kernels.

Kept identical for callers: ``LightGCN(num_users, num_items, num_layers=4, dim_h=64)``,
``forward(edge_index) -> (user_final, item_final)`` (views of one [N,64] tensor, differentiable
w.r.t. both embedding weights), ``get_embeddings(user_indices, item_indices)`` (raw layer-0 rows),
attributes ``num_users, num_items, num_layers, dim_h, user_embedding, item_embedding, convs`` and a
``state_dict()`` of exactly ``user_embedding.weight`` / ``item_embedding.weight`` so the reference's
``best_model.pth`` loads (utils/train_test.py:251,280).

What changed underneath: the K ``LGConv`` calls + stack/mean collapse into ONE fused K-layer
propagation over a cached CSR pair (``lgcn_propagate_fwd``), and autograd's gather/scatter chain
becomes the transpose propagation ``lgcn_propagate_bwd``.
"""
from __future__ import annotations

import warnings
from collections import OrderedDict
from ctypes import byref
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from .._lib import DIM, Graph, LgcnError, check, lib, require_cuda, stream_ptr


class GraphCache:
    """edge_index tensor -> Graph.  The reference recomputes the normalisation in every layer of
    every forward; its loaders hand the SAME tensors back every epoch
    (data/dataset_handler.py:277-285), so the CSR/degree build is done once per tensor.  The key
    holds a reference to the tensor (its storage cannot be recycled under us) and its version
    counter (in-place edits invalidate the entry)."""

    def __init__(self, capacity: int = 512):
        self.capacity = capacity
        self._d: "OrderedDict[int, tuple]" = OrderedDict()

    def get(self, edge_index: torch.Tensor, num_users: int, num_items: int) -> Graph:
        key = (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()))
        hit = self._d.get(key)
        if hit is not None and hit[0] is edge_index and hit[1] == edge_index._version \
                and hit[2].num_users == num_users and hit[2].num_items == num_items:
            self._d.move_to_end(key)
            return hit[2]
        g = Graph(edge_index, num_users, num_items)
        self._d[key] = (edge_index, edge_index._version, g)
        self._d.move_to_end(key)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
        return g

    def clear(self):
        self._d.clear()


class _Propagate(torch.autograd.Function):
    """final = sum_k A^k cat(U, I) / (K+1)^2 ; backward = the transpose propagation."""

    @staticmethod
    def forward(ctx, user_w, item_w, graph: Graph, num_layers: int):
        require_cuda(user_w, "user_embedding.weight", torch.float32)
        require_cuda(item_w, "item_embedding.weight", torch.float32)
        uw, iw = user_w.contiguous(), item_w.contiguous()
        n = graph.num_nodes
        final = torch.empty(n, DIM, dtype=torch.float32, device=uw.device)
        work = torch.empty(max(num_layers - 1, 1) * n * DIM, dtype=torch.float32, device=uw.device)
        check(lib().lgcn_propagate_fwd(graph.ref, uw.data_ptr(), iw.data_ptr(), num_layers, final.data_ptr(),
                                       None, work.data_ptr(), work.numel() * 4, stream_ptr(uw.device)))
        ctx.graph, ctx.num_layers = graph, num_layers
        return final

    @staticmethod
    def backward(ctx, grad_final):
        g, k = ctx.graph, ctx.num_layers
        gf = grad_final.contiguous()
        n = g.num_nodes
        grad = torch.empty(n, DIM, dtype=torch.float32, device=gf.device)
        work = torch.empty(2 * n * DIM, dtype=torch.float32, device=gf.device)
        check(lib().lgcn_propagate_bwd(g.ref, gf.data_ptr(), k, None, None, None, 0.0, grad.data_ptr(), None,
                                       work.data_ptr(), work.numel() * 4, stream_ptr(gf.device)))
        return grad[: g.num_users], grad[g.num_users:], None, None


class _Spmm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, graph: Graph):
        require_cuda(x, "x", torch.float32)
        xc = x.contiguous()
        out = torch.empty_like(xc)
        check(lib().lgcn_spmm(graph.ref, xc.data_ptr(), out.data_ptr(), 0, stream_ptr(xc.device)))
        ctx.graph = graph
        return out

    @staticmethod
    def backward(ctx, grad_out):
        go = grad_out.contiguous()
        gx = torch.empty_like(go)
        check(lib().lgcn_spmm(ctx.graph.ref, go.data_ptr(), gx.data_ptr(), 1, stream_ptr(go.device)))
        return gx, None


class LGConv(nn.Module):
    """Parameter-free stand-in for ``torch_geometric.nn.LGConv`` with the call the reference makes:
    ``conv(x=emb, edge_index=edge_index) -> [N, 64]`` (models/light_gcn.py:33), i.e. one
    symmetric-normalised propagation out[c] = sum_{r->c} deg(r)^-1/2 deg(c)^-1/2 x[r] with the
    in-degree of the CURRENT edge list.  ``LightGCN.forward`` does not loop over these; it runs
    the fused K-layer kernel."""

    def __init__(self, num_users: Optional[int] = None, cache: Optional[GraphCache] = None):
        super().__init__()
        self._num_users = num_users
        self._cache = cache if cache is not None else GraphCache(8)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        if x.dim() != 2 or x.size(1) != DIM:
            raise LgcnError(f"LGConv expects x of shape [N,{DIM}], got {tuple(x.shape)}")
        nu = self._num_users
        if nu is None:   # infer the user/movie boundary of a bipartite list: users are the smaller ids
            nu = int(torch.minimum(edge_index[0], edge_index[1]).max().item()) + 1 if edge_index.numel() else 0
        g = self._cache.get(edge_index, nu, x.size(0) - nu)
        return _Spmm.apply(x, g)


class LightGCN(nn.Module):
    def __init__(self, num_users, num_items, num_layers=4, dim_h=64):
        super().__init__()
        if dim_h != DIM:
            raise LgcnError(f"the sm_100a kernels are specialised for dim_h={DIM} (256-byte rows); got {dim_h}")
        if not 1 <= num_layers <= 4:
            raise LgcnError(f"num_layers must be in [1,4], got {num_layers}")
        self.num_users = num_users
        self.num_items = num_items
        self.num_layers = num_layers
        self.dim_h = dim_h

        self.user_embedding = nn.Embedding(num_embeddings=num_users, embedding_dim=dim_h)
        self.item_embedding = nn.Embedding(num_embeddings=num_items, embedding_dim=dim_h)
        self._graphs = GraphCache()
        # kept for API parity (`model.convs`); stateless, absent from state_dict like PyG's LGConv
        self.convs = nn.ModuleList(LGConv(num_users, self._graphs) for _ in range(num_layers))
        nn.init.normal_(self.user_embedding.weight, std=0.01)
        nn.init.normal_(self.item_embedding.weight, std=0.01)

    def graph(self, edge_index: torch.Tensor) -> Graph:
        """CSR pair + normalisation for ``edge_index`` (cached per tensor)."""
        return self._graphs.get(edge_index, self.num_users, self.num_items)

    def forward(self, edge_index) -> Tuple[torch.Tensor, torch.Tensor]:
        g = self.graph(edge_index)
        emb_final = _Propagate.apply(self.user_embedding.weight, self.item_embedding.weight, g, self.num_layers)
        return torch.split(emb_final, [self.num_users, self.num_items])

    def get_embeddings(self, user_indices=None, item_indices=None):
        """Raw layer-0 rows, never propagated (models/light_gcn.py:42-64)."""
        uw, iw = self.user_embedding.weight, self.item_embedding.weight
        if user_indices is not None and item_indices is not None:
            return uw[user_indices.to(uw.device)], iw[item_indices.to(iw.device)]
        if user_indices is not None:
            return uw[user_indices.to(uw.device)], None
        if item_indices is not None:
            return None, iw[item_indices.to(iw.device)]
        warnings.warn("Both indices not provided", UserWarning)
        return None, None
