"""TEST HELPER: float64 restatement of the full-graph training step on whatever device holds the inputs, chunked so
that the [E,64] / [P,64] temporaries stay bounded at BASELINE sizes.  Follows the reference's op sequence:
PyG gcn_norm + index_select / mul / scatter_add per layer (models/light_gcn.py:28-40), cosine BPR + regulariser
(utils/train_test.py:18-64); gradients are derived by hand (Horner transpose propagation), checked against autograd
in tests/test_gpu_at_scale.py at ML-25M size."""
import torch


def norm_weights(edge_index, n):
    row, col = edge_index[0], edge_index[1]
    deg = torch.bincount(col, minlength=n).double()
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    return row, col, dis


def spmm(row, col, dis, x, transpose=False, chunk=1 << 24):
    """out[c] = sum_{r->c} dis[r] dis[c] x[r]   (transpose: out[r] = sum dis[r] dis[c] x[c])"""
    out = torch.zeros_like(x)
    src, dst = (col, row) if transpose else (row, col)
    for b in range(0, row.numel(), chunk):
        s, d = src[b:b + chunk], dst[b:b + chunk]
        w = (dis[s] * dis[d])[:, None]
        out.index_add_(0, d, w * x.index_select(0, s))
    return out


def forward(edge_index, e0, k):
    n = e0.shape[0]
    row, col, dis = norm_weights(edge_index, n)
    x, acc = e0, e0.clone()
    for _ in range(k):
        x = spmm(row, col, dis, x)
        acc += x
    return acc / float((k + 1) ** 2)


def step_loss_and_grad(edge_index, e0, k, nu, neg, coeff=5e-3, chunk=1 << 21):
    """(loss, dL/de0, final) in float64; neg = item ids [P] for the user->movie edges in edge order."""
    n = e0.shape[0]
    row, col, dis = norm_weights(edge_index, n)
    final = forward(edge_index, e0, k)
    m = row < nu
    u, pos, ng = row[m], col[m], neg + nu
    p = u.numel()
    rn = 1.0 / final.norm(dim=1)
    G = torch.zeros_like(final)
    total = 0.0
    cnt = torch.zeros(n, dtype=torch.float64, device=e0.device)
    for b in range(0, p, chunk):
        s = slice(b, min(p, b + chunk))
        uu, pp, nn = u[s], pos[s], ng[s]
        uh, ph, nh = final[uu] * rn[uu, None], final[pp] * rn[pp, None], final[nn] * rn[nn, None]
        cp, cn = (uh * ph).sum(1), (uh * nh).sum(1)
        x = 10.0 * (cp - cn)
        total += float(torch.nn.functional.softplus(x).sum())
        sg = -torch.sigmoid(x) / p                                   # dL/dcos+ = -dL/dcos-
        G.index_add_(0, uu, (sg[:, None] * (ph - nh) - (sg * (cp - cn))[:, None] * uh) * rn[uu, None])
        G.index_add_(0, pp, (sg[:, None] * uh - (sg * cp)[:, None] * ph) * rn[pp, None])
        G.index_add_(0, nn, (-sg[:, None] * uh + (sg * cn)[:, None] * nh) * rn[nn, None])
        for idx in (uu, pp, nn):
            cnt.index_add_(0, idx, torch.ones(idx.numel(), dtype=torch.float64, device=e0.device))
    reg = float((cnt * e0.pow(2).sum(1)).sum())
    loss = -total / (10.0 * p) + coeff * reg / (64.0 * p)
    h = G.clone()
    for _ in range(k):
        h = G + spmm(row, col, dis, h, transpose=True)
    grad = h / float((k + 1) ** 2) + (2.0 * coeff / (64.0 * p)) * cnt[:, None] * e0
    return loss, grad, final
