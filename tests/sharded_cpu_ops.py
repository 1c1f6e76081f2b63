"""TEST DOUBLE (not product code): a torch-CPU implementation of the ``ops`` interface that
lgcn_b200.sharded.ShardedTrainer drives, so the sharded orchestration (row ownership, edge shards with global
triplet numbers, all-gathers, owner-computes BPR, owned-row Adam) can be exercised over gloo without a GPU.  Same
method contract as sharded.CudaOps -- every stage touches ONLY the rows this rank owns and reads remote rows only
from the exchanged tables, and the rank only looks at its SHARD of the edge list; math in float64, written from
the formulas in DESIGN.md (pre-scaled propagation, Horner backward, BPR-cosine gradients), independent of the
CUDA sources."""
import torch

DIM = 64


class TorchOps:
    def __init__(self, edge_index, num_users, num_items, num_layers, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=1.0, dtype=torch.float64):
        self.nu, self.ni, self.n, self.k = num_users, num_items, num_users + num_items, num_layers
        self.edge_index, self.dt = edge_index, dtype
        self.lr, self.b1, self.b2, self.eps, self.max_norm = lr, betas[0], betas[1], eps, max_norm
        self.p2p = False

    num_triplets = property(lambda self: self.P)

    def degrees(self):
        ei = self.edge_index
        return torch.bincount(ei[1], minlength=self.n), torch.bincount(ei[0], minlength=self.n)

    def bind(self, plan, rank):
        from lgcn_b200.sharded import EdgeShard
        n, dt, nu = self.n, self.dt, self.nu
        self.segs = plan.segments(rank)
        shard = EdgeShard.build(self.edge_index, plan, rank)
        self.P = shard.num_triplets
        del self.edge_index                       # from here on the rank sees its shard only
        row, col = shard.edges[0], shard.edges[1]
        own = torch.zeros(n, dtype=torch.bool)
        for rb, re in self.segs:
            own[rb:re] = True
        self.own = own
        # degrees / normalisation of OWNED rows are complete in the shard (the others are garbage and never used)
        self.in_deg = torch.bincount(col, minlength=n)
        self.out_deg = torch.bincount(row, minlength=n)
        deg = self.in_deg.to(dt)
        self.deg = deg
        self.dis = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
        ones = torch.ones(row.numel(), dtype=dt)
        self.A = torch.sparse_coo_tensor(torch.stack([col, row]), ones, (n, n)).coalesce()       # raw in-sum
        self.At = torch.sparse_coo_tensor(torch.stack([row, col]), ones, (n, n)).coalesce()      # raw out-sum
        um = row < nu
        t_glob = (torch.arange(int(um.sum())) if shard.trip_global is None else shard.trip_global.long())
        self.sh_user, self.sh_pos, self.sh_t = row[um], col[um], t_glob       # the shard's triplets, global numbers
        z = lambda *s: torch.zeros(*s, dtype=dt)
        self.y = [z(n, DIM) for _ in range(self.k)]
        self.z = [z(n, DIM) for _ in range(2)]
        self.zg = z(n, DIM)
        self.final, self.rnorm, self.G, self.grad = z(n, DIM), z(n), z(n, DIM), z(n, DIM)
        self.neg_count = torch.zeros(self.ni, dtype=torch.int32)
        self.accum = torch.zeros(4, dtype=torch.float64)
        self.loss = z(1)
        self.m, self.v = z(n, DIM), z(n, DIM)
        self.step = 0
        # triplet index: every rank fills the entries of its own users' triplets, the trainer sums them
        self.trip_user = torch.zeros(max(self.P, 1), dtype=torch.int32)
        self.trip_pos = torch.zeros(max(self.P, 1), dtype=torch.int32)
        mine = own[self.sh_user]
        self.trip_user[self.sh_t[mine]] = self.sh_user[mine].to(torch.int32)
        self.trip_pos[self.sh_t[mine]] = self.sh_pos[mine].to(torch.int32)

    def set_weights(self, uw, iw):
        self.uw, self.iw = uw, iw

    def _e0(self):
        return torch.cat([self.uw, self.iw])

    def _rows(self, full):
        """keep only owned rows of a freshly computed full-height result"""
        return [(rb, re, full[rb:re]) for rb, re in self.segs]

    def step_begin(self):
        self.step += 1
        self.accum.zero_()

    def prescale(self):
        for rb, re in self.segs:
            self.y[0][rb:re] = self.dis[rb:re, None] * self._e0()[rb:re]

    def fwd_layer(self, k, normalized=True):      # this double always keeps the raw rows (it re-normalises)
        full = self.A @ self.y[k - 1]
        for rb, re, raw in self._rows(full):
            deg, dis = self.deg[rb:re, None], self.dis[rb:re, None]
            if k < self.k:
                self.y[k][rb:re] = torch.where(deg > 0, raw / deg.clamp(min=1), torch.zeros_like(raw))
            else:
                s = sum((self.y[i][rb:re] for i in range(1, self.k)), torch.zeros_like(raw))
                f = (self._e0()[rb:re] + deg.sqrt() * s + dis * raw) / float((self.k + 1) ** 2)
                self.final[rb:re] = f
                self.rnorm[rb:re] = 1.0 / f.norm(dim=1)

    def bpr(self, neg):
        """Owner-computes: dL/dfinal rows of owned users from their triplets, of owned items from the triplets in which
        they are the positive (shard in-edges) or the sampled negative (ANY triplet: found through the exchanged
        triplet index)."""
        nu, P, F = self.nu, self.P, self.final
        nrm = lambda x: x / x.norm(dim=1, keepdim=True)
        (ub, ue), (ib, ie) = self.segs
        self.neg_count.zero_()
        self.G.zero_()

        def scalars(u, p, ng):
            uh, ph, nh = nrm(F[u]), nrm(F[p]), nrm(F[ng])
            cp, cn = (uh * ph).sum(1), (uh * nh).sum(1)
            x = 10 * (cp - cn)
            return uh, ph, nh, cp, cn, x, -torch.sigmoid(x)

        # user role
        m = self.own[self.sh_user]
        u, p, ng = self.sh_user[m], self.sh_pos[m], neg[self.sh_t[m]] + nu
        if u.numel():
            uh, ph, nh, cp, cn, x, s = scalars(u, p, ng)
            self.accum[0] += float(torch.nn.functional.softplus(x).sum())
            g = (s[:, None] * (ph - nh) - (s * (cp - cn))[:, None] * uh) / F[u].norm(dim=1, keepdim=True) / P
            self.G.index_add_(0, u, g)
        # positive role
        m = self.own[self.sh_pos]
        u, p, ng = self.sh_user[m], self.sh_pos[m], neg[self.sh_t[m]] + nu
        if u.numel():
            uh, ph, nh, cp, cn, x, s = scalars(u, p, ng)
            self.G.index_add_(0, p, (s[:, None] * uh - (s * cp)[:, None] * ph) / F[p].norm(dim=1, keepdim=True) / P)
        # negative role
        t = torch.nonzero((neg + nu >= ib) & (neg + nu < ie)).flatten()
        self.neg_count += torch.bincount(neg[t], minlength=self.ni).to(torch.int32)
        if t.numel():
            u, p, ng = self.trip_user[t].long(), self.trip_pos[t].long(), neg[t] + nu
            uh, ph, nh, cp, cn, x, s = scalars(u, p, ng)
            self.G.index_add_(0, ng, (-s[:, None] * uh + (s * cn)[:, None] * nh) / F[ng].norm(dim=1, keepdim=True) / P)
        for rb, re in self.segs:
            self.zg[rb:re] = self.dis[rb:re, None] * self.G[rb:re]

    def bwd_layer(self, j, coeff):
        zin = self.zg if j == 1 else self.z[j & 1]
        full = self.At @ zin
        for rb, re, S in self._rows(full):
            h = self.G[rb:re] + self.dis[rb:re, None] * S
            if j < self.k:
                self.z[(j - 1) & 1][rb:re] = self.dis[rb:re, None] * h
                continue
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

    def clip_adam(self, coeff):
        clip = min(1.0, self.max_norm / (float(self.accum[2].sqrt()) + 1e-6))
        for rb, re in self.segs:
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
