"""numpy stand-in for the kernel backend of ppanggolin_b200.nem.sharded.ShardedEM -- TEST HELPER ONLY.

It exists so that the exchange logic of the row-sharded driver (row partition, all-reduce of the M-step statistics,
all-gather of the log-densities, replicated sweep) can run under gloo with world_size 2 on a CPU-only machine.  The
product never imports it; the product backend is CudaBackend (the C ABI)."""
import numpy as np
import torch

import npref


class NumpyBackend:
    def __init__(self, X_local, D, N_total, K, free_disp, graph, logspace=True):
        self.X = X_local.astype(np.uint8)
        self.N, self.K, self.D, self.fd = N_total, K, D, free_disp
        self.rowptr, self.col, self.w = graph
        self.logf_full = torch.zeros((N_total, K), dtype=torch.float64)
        self.T = [torch.zeros((N_total, K), dtype=torch.float32) for _ in range(2)]
        self._stats = [torch.zeros((K, D), dtype=torch.int32), torch.zeros((K, D), dtype=torch.float64),
                       torch.zeros(K, dtype=torch.int32), torch.zeros(K, dtype=torch.float64)]
        self.maxdiff, self.status = 0.0, 0
        self.launches = 0

    def init_params(self):
        self.pk, self.mu, self.eps = npref.init_params(self.K, self.D)

    def accumulate(self, T_local):
        t = T_local.numpy()
        K = self.K
        onehot = ((t == 1.0).sum(1) == 1) & ((t == 0.0).sum(1) == K - 1)
        cls = t.argmax(1)
        cnt, fsum, nkc, nkf = self._stats
        for k in range(K):
            sel = onehot & (cls == k)
            cnt[k] = torch.from_numpy(self.X[sel].sum(0).astype(np.int32))
            nkc[k] = int(sel.sum())
        fz = ~onehot
        fsum[:] = torch.from_numpy(t[fz].astype(np.float64).T @ self.X[fz].astype(np.float64))
        nkf[:] = torch.from_numpy(t[fz].astype(np.float64).sum(0))

    def stats(self):
        return self._stats

    def finalize(self):
        cnt, fsum, nkc, nkf = [s.numpy() for s in self._stats]
        nk = nkc + nkf
        if (nk <= 1e-20).any():
            self.status = 1
            return
        s1 = cnt + fsum
        s0 = nk[:, None] - s1
        w0 = s0.astype(np.float32); half = (nk.astype(np.float32) * np.float32(0.5))[:, None]
        self.mu = np.where(w0 < half, 1.0, np.where(w0 > half, 0.0, 0.5)).astype(np.float32)
        iner = np.where(w0 < half, s0, np.where(w0 > half, s1, 0.5 * nk[:, None]))
        if self.fd:
            self.eps = (iner / nk[:, None]).astype(np.float32)
        else:
            self.eps = np.repeat((iner.sum(1) / (self.D * nk)).astype(np.float32)[:, None], self.D, 1)
        self.pk = (nk / self.N).astype(np.float32)

    def density(self, logf_local):
        logf_local[:] = torch.from_numpy(npref.logf(self.X, self.mu, self.eps))

    def sweep(self, T_old, T_new, beta_now):
        lf = self.logf_full.numpy()
        old = T_old.numpy()
        new = T_new.numpy()
        logpk = np.log(self.pk.astype(np.float64))
        beta = float(np.float32(beta_now))
        for i in range(self.N):
            ctx = np.zeros(self.K, np.float32)
            if beta != 0.0:
                for n in range(self.rowptr[i], self.rowptr[i + 1]):
                    j = self.col[n]
                    src = new if j < i else old
                    ctx = (ctx + np.float32(self.w[n]) * src[j]).astype(np.float32)
            a = logpk + lf[i] + beta * ctx.astype(np.float64)
            m = a.max()
            e = np.exp(a - m)
            new[i] = (e / e.sum()).astype(np.float32)
        self.maxdiff = float(np.abs(new - old).max())

    def read_flags(self):
        return self.maxdiff, self.status

    def criteria(self, T, beta):
        return [0.0] * 5
