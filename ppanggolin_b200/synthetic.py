"""Synthetic pangenomes of the shapes BASELINE.json names.

One seed -> (presence matrix, neighbour graph) in the array form the NEM path consumes.  Two generators:

* :func:`synth_pangenome` (numpy, exact graph semantics): every genome is one circular contig listing its present
  families in a block-shuffled order; edges and weights then follow ``graph/makeGraph.py:114-136`` +
  ``nem/partition.py:393-433`` (pair count / #genomes, ``round(.,4)``, ``sm_degree`` cut).  For N*D up to ~1e8.
* :func:`synth_pangenome_large` (torch, on the GPU): bit-packed matrix generated directly in HBM plus a backbone
  graph with the same degree / weight statistics.  For the 50 000 x 200 000 bench shape (1.25 GB of bits).

Bit packing everywhere in this package: row-major uint32 words, bit ``b`` of word ``w`` of a row is genome column
``32*w + b``; rows are padded with zero bits to a multiple of 128 columns (16-byte vector loads).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def row_stride_words(D: int) -> int:
    """Words per matrix row.  The C ABI needs a multiple of 4 (16-byte rows); wide matrices are padded to whole 128-byte
    lines: a row that starts inside a 32-byte sector doubles the L2 sector requests of the 16-byte cp.async copies the
    streaming kernels use (measured with ncu: 77 M sectors instead of 39 M on the 50 000-genome matrix)."""
    words = (D + 31) // 32
    return (words + 3) // 4 * 4 if words <= 64 else (words + 31) // 32 * 32


def pack_rows(X: np.ndarray) -> np.ndarray:
    """(N, D) 0/1 -> (N, row_stride_words(D)) uint32."""
    X = np.ascontiguousarray(X, dtype=np.uint8)
    N, D = X.shape
    W = row_stride_words(D)
    padded = np.zeros((N, W * 32), np.uint8)
    padded[:, :D] = X
    return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(N, W)


def unpack_rows(bits: np.ndarray, D: int) -> np.ndarray:
    b = np.ascontiguousarray(bits).view(np.uint8)
    return np.unpackbits(b, axis=1, bitorder="little")[:, :D]


@dataclass
class SynthPangenome:
    X: np.ndarray          # (N, D) uint8 (None for the large generator)
    bits: object           # (N, W) uint32 numpy array or int32 torch tensor on the GPU
    rowptr: np.ndarray     # (N+1,) int32 neighbour lists as NEM keeps them
    col: np.ndarray        # (nnz,) int32
    w: np.ndarray          # (nnz,) float32 (values are round(x, 4))
    edges_weight: float    # sum of kept edge weights / 2 (partition.py:416,436)
    true_class: np.ndarray # 0 persistent / 1 shell / 2 cloud, generator ground truth
    N: int
    D: int

    def beta_eff(self, beta: float = 2.5) -> float:
        """partition.py:872 -- beta * nb_fam / edges_weight."""
        return beta * (self.N / self.edges_weight)


def _family_probs(N, rng):
    u = rng.random(N)
    cls = np.where(u < 0.40, 0, np.where(u < 0.55, 1, 2)).astype(np.int8)
    p = np.empty(N)
    p[cls == 0] = 0.98
    p[cls == 1] = rng.uniform(0.15, 0.85, int((cls == 1).sum()))
    p[cls == 2] = rng.uniform(0.01, 0.02, int((cls == 2).sum()))
    return cls, p


def synth_matrix(N: int, D: int, seed: int = 42):
    rng = np.random.default_rng(seed)
    cls, p = _family_probs(N, rng)
    X = (rng.random((N, D)) < p[:, None]).astype(np.uint8)
    empty = np.flatnonzero(X.sum(1) == 0)
    X[empty, rng.integers(0, D, len(empty))] = 1
    return X, cls, p, rng


def graph_from_contigs(X: np.ndarray, rng, sm_degree: float = 10, chain_order: bool = False):
    """Neighbour lists with partition.py:393-433 semantics from one circular contig per genome."""
    N, D = X.shape
    # backbone position of each family: random, so that the row index order (the Gauss-Seidel visit order) is unrelated
    # to the genomic order; chain_order=True is the adversarial variant where rows follow the backbone.
    backbone = np.arange(N) if chain_order else rng.permutation(N)
    # blocks of 10-50 consecutive backbone positions
    cuts = [0]
    while cuts[-1] < N:
        cuts.append(min(N, cuts[-1] + int(rng.integers(10, 51))))
    nblk = len(cuts) - 1
    keys = []
    for g in range(D):
        order = np.arange(nblk)
        for _ in range(int(rng.integers(0, 4))):  # a few block rearrangements per genome
            a, b = rng.integers(0, nblk, 2)
            order[a], order[b] = order[b], order[a]
        pos = np.concatenate([np.arange(cuts[b], cuts[b + 1]) for b in order])
        fams = backbone[pos]
        fams = fams[X[fams, g] != 0]
        if len(fams) < 2:
            continue
        a, b = fams, np.roll(fams, -1)  # consecutive genes + circular closure (makeGraph.py:114-136)
        if len(fams) == 2:
            a, b = a[:1], b[:1]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        keys.append(lo.astype(np.int64) * N + hi)
    keys = np.concatenate(keys) if keys else np.zeros(0, np.int64)
    uniq, cnt = np.unique(keys, return_counts=True)
    src, dst = (uniq // N).astype(np.int64), (uniq % N).astype(np.int64)
    # directed entries, both directions; per-row order = order of edge creation is irrelevant to the arrays' consumers
    r = np.concatenate([src, dst]); c = np.concatenate([dst, src]); n = np.concatenate([cnt, cnt])
    o = np.argsort(r, kind="stable")
    r, c, n = r[o], c[o], n[o]
    dist = n / D                                   # partition.py:408
    wr = np.round(dist, 4)                         # partition.py:413
    deg = np.bincount(r, minlength=N)
    keep_row = (deg > 0) & (deg < sm_degree)       # partition.py:415
    sum_dist = np.bincount(r, weights=dist, minlength=N)
    edges_weight = float(sum_dist[keep_row].sum()) / 2  # partition.py:416,436
    starts = np.concatenate([[0], np.cumsum(deg)])
    rowptr = np.zeros(N + 1, np.int64); cols = []; ws = []
    for i in np.flatnonzero(keep_row):
        a, b = starts[i], starts[i + 1]
        ci, wi = c[a:b], wr[a:b]
        nz = wi != 0.0                             # nem_exe.c:1413-1451: zero weights dropped AFTER indices were stored
        nv = int(nz.sum())
        cols.append(ci[:nv]); ws.append(wi[nz])
        rowptr[i + 1] = nv
    rowptr = np.cumsum(rowptr).astype(np.int32)
    col = (np.concatenate(cols) if cols else np.zeros(0)).astype(np.int32)
    w = (np.concatenate(ws) if ws else np.zeros(0)).astype(np.float32)
    return rowptr, col, w, edges_weight


def synth_pangenome(N: int, D: int, seed: int = 42, sm_degree: float = 10, chain_order: bool = False) -> SynthPangenome:
    X, cls, _, rng = synth_matrix(N, D, seed)
    rowptr, col, w, ew = graph_from_contigs(X, rng, sm_degree, chain_order)
    return SynthPangenome(X, pack_rows(X), rowptr, col, w, ew, cls, N, D)


def backbone_graph(N: int, p: np.ndarray, rng, sm_degree: float = 10):
    """Cheap graph with contig-walk statistics for shapes where walking D genomes is too slow: each family is linked to
    its 2 backbone neighbours on either side with weight ~ joint presence; 2 % of families are hubs (degree >= sm_degree,
    so they lose their own list but stay in others', partition.py:415-433)."""
    backbone = rng.permutation(N)
    pos_of = np.empty(N, np.int64); pos_of[backbone] = np.arange(N)
    src = []; dst = []; wt = []
    for off in (1, 2):
        a = backbone; b = np.roll(backbone, -off)
        ww = p[a] * p[b] * (1.0 if off == 1 else 0.15)
        src.append(a); dst.append(b); wt.append(ww)
    hubs = rng.choice(N, max(1, N // 50), replace=False)
    for h in hubs:
        other = rng.choice(N, 12, replace=False)
        other = other[other != h]
        src.append(np.full(len(other), h)); dst.append(other); wt.append(p[h] * p[other] * 0.05)
    src = np.concatenate(src); dst = np.concatenate(dst); wt = np.concatenate(wt)
    lo, hi = np.minimum(src, dst), np.maximum(src, dst)
    key, first = np.unique(lo * N + hi, return_index=True)
    lo, hi, wt = lo[first], hi[first], wt[first]
    r = np.concatenate([lo, hi]); c = np.concatenate([hi, lo]); ww = np.concatenate([wt, wt])
    o = np.argsort(r, kind="stable"); r, c, ww = r[o], c[o], ww[o]
    wr = np.round(ww, 4)
    deg = np.bincount(r, minlength=N)
    keep_row = (deg > 0) & (deg < sm_degree)
    keep = keep_row[r] & (wr != 0.0)
    sum_dist = np.bincount(r, weights=ww, minlength=N)
    edges_weight = float(sum_dist[keep_row].sum()) / 2
    r, c, wr = r[keep], c[keep], wr[keep]
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=N))]).astype(np.int32)
    return rowptr, c.astype(np.int32), wr.astype(np.float32), edges_weight


def synth_pangenome_large(N: int, D: int, seed: int = 42, device="cuda", rows=None, rows_per_block: int = 2048) -> SynthPangenome:
    """Bit-packed matrix generated block by block on ``device`` (torch); graph from :func:`backbone_graph`.

    ``rows=(lo, hi)`` generates only that row range (one rank's shard of a row-sharded run); blocks are seeded by their
    absolute position, so the union of the shards is the same matrix for every world size."""
    import torch

    rng = np.random.default_rng(seed)
    cls, p = _family_probs(N, rng)
    lo, hi = (0, N) if rows is None else rows
    W = row_stride_words(D)
    gen = torch.Generator(device=device)
    bits = torch.empty((hi - lo, W), dtype=torch.int32, device=device)
    shifts = torch.arange(32, device=device, dtype=torch.int64)
    pt = torch.from_numpy(p.astype(np.float32)).to(device)
    b0 = (lo // rows_per_block) * rows_per_block
    for r0 in range(b0, hi, rows_per_block):
        r1 = min(N, r0 + rows_per_block)
        gen.manual_seed(seed * 1_000_003 + r0)
        u = torch.rand((r1 - r0, W * 32), device=device, generator=gen)
        x = (u < pt[r0:r1, None])
        del u
        x[:, D:] = False
        none = ~x.any(1)
        x[none, 0] = True  # all-zero families get one gene (they would not be in the matrix, partition.py:384)
        words = (x.view(r1 - r0, W, 32).to(torch.int64) << shifts).sum(-1)
        words = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
        a, b = max(r0, lo), min(r1, hi)
        bits[a - lo:b - lo] = words[a - r0:b - r0]
        del x, words
    rowptr, col, w, ew = backbone_graph(N, p, rng)
    return SynthPangenome(None, bits, rowptr, col, w, ew, cls, N, D)


def synth_pangenome_blocked(N: int, D: int, seed: int = 42, block: int = 1024, keep_X: bool = False) -> SynthPangenome:
    """CPU generator for wide matrices (numpy only, so the same seed gives the same instance wherever it runs): the matrix
    is drawn in row blocks seeded by their first row, so the float randoms of one block (block x D float32) are all that
    is ever held; graph from :func:`backbone_graph`.  ``keep_X`` also returns the unpacked uint8 matrix (oracle input)."""
    rng = np.random.default_rng(seed)
    cls, p = _family_probs(N, rng)
    W = row_stride_words(D)
    bits = np.zeros((N, W), np.uint32)
    X = np.zeros((N, D), np.uint8) if keep_X else None
    for r0 in range(0, N, block):
        r1 = min(N, r0 + block)
        rb = np.random.default_rng([seed, r0])
        xb = (rb.random((r1 - r0, D), dtype=np.float32) < p[r0:r1, None].astype(np.float32)).astype(np.uint8)
        xb[xb.sum(1) == 0, 0] = 1
        bits[r0:r1] = pack_rows(xb)
        if keep_X:
            X[r0:r1] = xb
    rowptr, col, w, ew = backbone_graph(N, p, rng)
    return SynthPangenome(X, bits, rowptr, col, w, ew, cls, N, D)
