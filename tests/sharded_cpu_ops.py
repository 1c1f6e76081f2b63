"""TEST DOUBLE (not product code): a torch-CPU implementation of the ``ops`` interface that
lgcn_b200.sharded.ShardedTrainer drives, so the sharded orchestration (row ownership, all-gathers,
all-reduces, owned-row Adam) can be exercised over gloo without a GPU.  Same method contract as
sharded.CudaOps; math in float64, written from the formulas in DESIGN.md (pre-scaled propagation,
Horner backward, BPR-cosine gradients), independent of the CUDA sources."""
import torch

DIM = 64


class TorchOps:
    def __init__(self, edge_index, num_users, num_items, num_layers, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=1.0, dtype=torch.float64):
        self.nu, self.ni, self.n, self.k = num_users, num_items, num_users + num_items, num_layers
        n, dt = self.n, dtype
        row, col = edge_index[0], edge_index[1]
        self.row, self.col = row, col
        self.in_deg = torch.bincount(col, minlength=n)
        self.out_deg = torch.bincount(row, minlength=n)
        deg = self.in_deg.to(dt)
        self.deg = deg
        self.dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
        ones = torch.ones(row.numel(), dtype=dt)
        self.A = torch.sparse_coo_tensor(torch.stack([col, row]), ones, (n, n)).coalesce()       # raw in-sum
        self.At = torch.sparse_coo_tensor(torch.stack([row, col]), ones, (n, n)).coalesce()      # raw out-sum
        um = row < num_users
        self.t_user, self.t_pos = row[um], col[um]
        self.P = int(um.sum())
        z = lambda *s: torch.zeros(*s, dtype=dt)
        self.y = [z(n, DIM) for _ in range(num_layers)]
        self.z = [z(n, DIM) for _ in range(min(2, max(num_layers - 1, 0)))]
        self.final, self.rnorm, self.G, self.grad = z(n, DIM), z(n), z(n, DIM), z(n, DIM)
        self.neg_count = torch.zeros(num_items, dtype=torch.int32)
        self.accum = torch.zeros(4, dtype=torch.float64)
        self.loss = z(1)
        self.m, self.v = z(n, DIM), z(n, DIM)
        self.step = 0
        self.lr, self.b1, self.b2, self.eps, self.max_norm = lr, betas[0], betas[1], eps, max_norm

    num_triplets = property(lambda self: self.P)

    def degrees(self):
        return self.in_deg, self.out_deg

    def set_weights(self, uw, iw):
        self.uw, self.iw = uw, iw

    def _e0(self):
        return torch.cat([self.uw, self.iw])

    def step_begin(self):
        self.step += 1
        self.accum.zero_()

    def prescale(self, rb, re):
        self.y[0][rb:re] = self.dis[rb:re, None] * self._e0()[rb:re]

    def fwd_layer(self, k, rb, re):
        raw = (self.A @ self.y[k - 1])[rb:re]
        deg, dis = self.deg[rb:re, None], self.dis[rb:re, None]
        if k < self.k:
            self.y[k][rb:re] = torch.where(deg > 0, raw / deg.clamp(min=1), torch.zeros_like(raw))
        else:
            s = sum((self.y[i][rb:re] for i in range(1, self.k)), torch.zeros_like(raw))
            f = (self._e0()[rb:re] + deg.sqrt() * s + dis * raw) / float((self.k + 1) ** 2)
            self.final[rb:re] = f
            self.rnorm[rb:re] = 1.0 / f.norm(dim=1)

    def bpr(self, neg, urb, ure):
        self.G.zero_()
        self.neg_count.zero_()
        own = (self.t_user >= urb) & (self.t_user < ure)
        u, p, ng = self.t_user[own], self.t_pos[own], neg[own] + self.nu
        self.neg_count += torch.bincount(ng - self.nu, minlength=self.ni).to(torch.int32)
        if u.numel() == 0:
            return
        F = self.final.clone().requires_grad_(True)
        nrm = lambda x: x / x.norm(dim=1, keepdim=True)
        cp = (nrm(F[u]) * nrm(F[p])).sum(1)
        cn = (nrm(F[u]) * nrm(F[ng])).sum(1)
        sp = torch.nn.functional.softplus(10 * (cp - cn))
        (-(sp.sum()) / (10.0 * self.P)).backward()
        self.G += F.grad
        self.accum[0] += float(sp.sum())

    def bwd_layer(self, j, rb, re, coeff):
        zin = self.dis[:, None] * self.G if j == 1 else self.z[j & 1]
        S = (self.At @ zin)[rb:re]
        h = self.G[rb:re] + self.dis[rb:re, None] * S
        if j < self.k:
            self.z[(j - 1) & 1][rb:re] = self.dis[rb:re, None] * h
            return
        cnt = torch.zeros(self.n, dtype=h.dtype)
        cnt[: self.nu] = self.out_deg[: self.nu].to(h.dtype)
        cnt[self.nu:] = self.in_deg[self.nu:].to(h.dtype) + self.neg_count.to(h.dtype)
        e0 = self._e0()[rb:re]
        g = h / float((self.k + 1) ** 2) + (2.0 * coeff / (64.0 * self.P)) * cnt[rb:re, None] * e0
        self.grad[rb:re] = g
        self.accum[1] += float((cnt[rb:re] * e0.pow(2).sum(1)).sum())
        self.accum[2] += float(g.pow(2).sum())

    def zbuf(self, j):
        return self.z[(j - 1) & 1]

    def clip_adam(self, rb, re, coeff):
        clip = min(1.0, self.max_norm / (float(self.accum[2].sqrt()) + 1e-6))
        g = self.grad[rb:re] * clip
        self.m[rb:re] = self.m[rb:re] + (1 - self.b1) * (g - self.m[rb:re])
        self.v[rb:re] = self.v[rb:re] * self.b2 + (1 - self.b2) * g * g
        bc1, bc2 = 1 - self.b1 ** self.step, 1 - self.b2 ** self.step
        upd = (self.lr / bc1) * self.m[rb:re] / (self.v[rb:re].sqrt() / bc2 ** 0.5 + self.eps)
        for w, lo in ((self.uw, 0), (self.iw, self.nu)):
            a, b = max(rb, lo), min(re, lo + w.shape[0])
            if a < b:
                w[a - lo:b - lo] -= upd[a - rb:b - rb]
        self.loss[0] = -self.accum[0] / (10.0 * self.P) + coeff * self.accum[1] / (64.0 * self.P)
