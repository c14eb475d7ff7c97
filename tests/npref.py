"""numpy closed forms of the individual kernels (test helpers; not product code)."""
import numpy as np


def init_params(K, D):
    import math
    base = np.float32(float(format(1.0 / K, ".12g")))
    pk = np.full(K, base, np.float32)
    pK = np.float32(1.0)
    for _ in range(K - 1):
        pK = np.float32(pK - base)
    pk[K - 1] = pK
    step = 0.5 / math.ceil(K / 2)
    pich = 0.1 if K == 2 else 0
    mu = np.zeros((K, D), np.float32); eps = np.zeros((K, D), np.float32)
    for k in range(1, K + 1):
        if k <= K / 2:
            mu[k - 1] = 1; eps[k - 1] = np.float32(step * k - pich)
        else:
            eps[k - 1] = np.float32(step * (K - k + 1) - pich)
    return pk, mu, eps


def logf(X, mu, eps):
    """log f_k(x_i) in float64 with the float32 ratio of nem_mod.c:660."""
    X = X.astype(np.float64)
    K, D = mu.shape
    out = np.zeros((X.shape[0], K))
    for k in range(K):
        e = eps[k].astype(np.float32)
        pos = e > 1e-20
        with np.errstate(divide="ignore"):
            L = np.where(pos, np.log(((np.float32(1) - e) / np.where(pos, e, 1)).astype(np.float64)), np.inf)
            l1 = np.where(pos, np.log((np.float32(1) - e).astype(np.float64)), 0.0)
        mism = (X != mu[k][None, :]) & (mu[k][None, :] != 0.5)
        with np.errstate(invalid="ignore"):
            s = np.where(mism, L[None, :], 0.0).sum(1)
        out[:, k] = l1.sum() - s
    return out


def softmax_ref_epilogue(lf, pk):
    """beta = 0 posterior with the reference arithmetic (float32 logf, double exp, S == 0 -> 1/K)."""
    K = lf.shape[1]
    f = np.exp(lf.astype(np.float32).astype(np.float64))
    num = pk.astype(np.float64)[None, :] * f
    S = num.sum(1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        T = np.where(S > 0, num / S, 1.0 / K)
    return T.astype(np.float32)


def mstep(X, T, free_disp):
    """fp64 M-step: centres by weighted median of a binary column, inertia, dispersion, proportions."""
    X = X.astype(np.float64); Td = T.astype(np.float64)
    N, D = X.shape; K = T.shape[1]
    nk = Td.sum(0)
    s1 = Td.T @ X
    s0 = nk[:, None] - s1
    w0 = s0.astype(np.float32); half = (nk.astype(np.float32) * np.float32(0.5))[:, None]   # float32-resolution comparison
    mu = np.where(w0 < half, 1.0, np.where(w0 > half, 0.0, 0.5)).astype(np.float32)
    iner = np.where(w0 < half, s0, np.where(w0 > half, s1, 0.5 * nk[:, None]))
    if free_disp:
        eps = (iner / nk[:, None]).astype(np.float32)
    else:
        eps = np.repeat((iner.sum(1) / (D * nk)).astype(np.float32)[:, None], D, 1)
    pk = (nk / N).astype(np.float32)
    return mu, eps, pk, s1, nk


def apply_votes_ref(codes, votes, validated, n_org, chunk_size):
    """validate_family (partition.py:784-802) for a sequence of chunk label vectors ``codes`` [n_chunks, F] (0 = P, 1 = S,
    2 = C, -1 = not in the chunk), chunk after chunk; ``votes`` [F, 4] (P, S, C, U) and ``validated`` [F] are updated in place."""
    for row in codes:
        for f in np.flatnonzero(row >= 0):
            votes[f, row[f]] += 1
            s_ = votes[f].sum()
            m_ = votes[f].max()
            if (s_ > n_org / chunk_size and m_ >= s_ * 0.5) or s_ > n_org:
                if not validated[f]:
                    if m_ < s_ * 0.5:
                        votes[f, 3] = n_org
                    validated[f] = True
