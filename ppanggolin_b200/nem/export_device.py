"""Device-resident array form of a pangenome and the CUDA exporter that derives NEM instances from it.

The reference walks the Python objects again for every genome subset it partitions -- the whole pangenome, the
K-selection sample, every chunk of the chunked mode -- and writes text files (``write_nem_input_files``,
/root/reference/ppanggolin/nem/partition.py:344-436; ~1.2 us per matrix cell).  Here the pangenome is described ONCE by
arrays in HBM (:class:`DevicePangenome`):

* ``presence``  bit-packed families x genomes matrix,
* the adjacency of every family in ``fam.edges`` order (other family, edge id),
* gene pairs per (edge, genome) as a CSR table over edges -- or, for synthetic pangenomes at scales where that table does
  not fit, the co-presence model (one gene pair in genome g iff both families of the edge are present in g),

and ``nemb200_export_samples`` (csrc/nemb200_export.cuh) derives the instances of a whole batch of genome subsets on the
device: column gather of the bit matrix, row compaction, edge coverage, weights ``round(pairs / n, 4)``, the max-degree
cut, what the C reader keeps of zero weights (NEM/nem_exe.c:1395-1451), ``sum of kept weights / 2`` and the Gauss-Seidel
level schedule -- no host round trip of any array.  ``tests/test_export_device.py`` checks every output against
``export.export_from_pangenome`` (the traversal that follows the reference line by line).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from ..synthetic import row_stride_words
from . import engine

_vp, _i64, _i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32


class _Pan(ctypes.Structure):           # struct nemb200_pangenome
    _fields_ = [("F", _i64), ("G", _i64), ("E", _i64), ("A", _i64), ("Wg", _i64), ("presence", _vp), ("adj_ptr", _vp),
                ("adj_other", _vp), ("adj_edge", _vp), ("cov_ptr", _vp), ("cov_org", _vp), ("cov_cnt", _vp), ("edge_fam", _vp)]


class _Sample(ctypes.Structure):        # struct nemb200_sample
    _fields_ = [("sel", _vp), ("n_sel", _i32), ("row_stride_words", _i64), ("bits", _vp), ("fam_rows", _vp), ("rowptr", _vp),
                ("col", _vp), ("w", _vp), ("level_ptr", _vp), ("level_rows", _vp), ("N", _i64), ("nnz", _i64),
                ("n_levels", _i32), ("edges_weight", ctypes.c_double)]


@dataclass
class DeviceInstance:
    """One exported NEM instance, resident in HBM."""
    inp: engine.DeviceInput       # matrix, CSR, level schedule
    fam_rows: object              # int32 [N]: family index of every instance row
    N: int
    D: int
    edges_weight: float

    def beta_eff(self, beta: float) -> float:
        """partition.py:324,872: beta * nb_fam / edges_weight (ZeroDivisionError without weighted edges, like the reference)."""
        return beta * (self.N / self.edges_weight)


def python_round4_lut(n_sel: int, max_pairs: int) -> np.ndarray:
    """lut[c] = float32 of Python's ``round(c / n_sel, 4)`` after its trip through text (partition.py:413 writes
    ``str(round(coverage / len(organisms), 4))``, nem_exe.c reads it with ``fscanf("%f")``)."""
    return np.array([np.float32(float(str(round(c / n_sel, 4)))) for c in range(max_pairs + 1)], np.float32)


class DevicePangenome:
    def __init__(self, presence, G: int, adj_ptr, adj_other, adj_edge, n_edges: int, cov=None, edge_fam=None,
                 max_pairs: Optional[int] = None, fam_names: Optional[List[str]] = None, org_names: Optional[List[str]] = None,
                 org_index: Optional[dict] = None):
        """All arrays are torch int32 tensors on one CUDA device.  ``cov`` = (cov_ptr [E+1], cov_org [M], cov_cnt [M]) or
        None with ``edge_fam`` [E, 2] (co-presence model)."""
        torch = engine._require_cuda()
        self.torch = torch
        self.presence = presence.contiguous()
        self.device = presence.device
        self.F, self.Wg = int(presence.shape[0]), int(presence.shape[1])
        self.G = int(G)
        self.adj_ptr, self.adj_other, self.adj_edge = adj_ptr.contiguous(), adj_other.contiguous(), adj_edge.contiguous()
        self.A, self.E = int(adj_other.numel()), int(n_edges)
        self.cov = tuple(t.contiguous() for t in cov) if cov is not None else None
        self.edge_fam = edge_fam.contiguous() if edge_fam is not None else None
        if self.E > 0 and self.cov is None and self.edge_fam is None:
            raise engine.NemB200Error("a pangenome with edges needs a coverage table or the edge families (co-presence model)")
        if max_pairs is None:
            if self.cov is not None and self.cov[2].numel():
                per_edge = torch.zeros(max(self.E, 1), dtype=torch.int64, device=self.device)
                ptr = self.cov[0].to(torch.int64)
                owner = torch.repeat_interleave(torch.arange(self.E, device=self.device), ptr[1:] - ptr[:-1])
                per_edge.index_add_(0, owner, self.cov[2].to(torch.int64))
                max_pairs = int(per_edge.max().item())
            else:
                max_pairs = self.G
        self.max_pairs = int(max_pairs)
        self.fam_names = fam_names if fam_names is not None else [f"fam{i}" for i in range(self.F)]
        self.org_names = org_names if org_names is not None else [f"org{i}" for i in range(self.G)]
        self.org_index = org_index
        self._luts = {}

    # ------------------------------------------------------------------------------------------------- constructors
    @staticmethod
    def from_arrays(arrays, device="cuda") -> "DevicePangenome":
        """From :class:`export_arrays.PangenomeArrays` (the one traversal of the Python objects)."""
        import torch
        from ..synthetic import pack_rows
        pres = arrays.presence.cpu().numpy().astype(np.uint8)
        F, G = pres.shape
        Wg = row_stride_words(G)
        bits = torch.from_numpy(pack_rows(pres).view(np.int32).reshape(F, Wg)).to(device)
        i32 = lambda t: t.to(device=device, dtype=torch.int32)
        # (edge, genome, pairs) COO -> CSR over edges
        ce, co, cc = arrays.cov_edge.cpu(), arrays.cov_org.cpu(), arrays.cov_cnt.cpu()
        order = torch.argsort(ce, stable=True)
        ce, co, cc = ce[order], co[order], cc[order]
        ptr = torch.zeros(arrays.n_edges + 1, dtype=torch.int64)
        if ce.numel():
            ptr[1:] = torch.cumsum(torch.bincount(ce, minlength=arrays.n_edges), 0)
        return DevicePangenome(bits, G, i32(arrays.adj_ptr), i32(arrays.adj_other), i32(arrays.adj_edge), arrays.n_edges,
                               cov=(i32(ptr), i32(co), i32(cc)), fam_names=list(arrays.fam_names), org_names=list(arrays.org_names),
                               org_index=arrays.org_index)

    @staticmethod
    def from_pangenome(pan, device="cuda") -> "DevicePangenome":
        """One traversal of the PPanGGOLiN objects (pangenome.py:213,323; geneFamily.py:272,296; edge.py:75)."""
        from .export_arrays import PangenomeArrays
        return DevicePangenome.from_arrays(PangenomeArrays.from_pangenome(pan, "cpu"), device)

    @staticmethod
    def synthetic_copresence(bits, G: int, edges: np.ndarray, device=None) -> "DevicePangenome":
        """Synthetic pangenome at scale: ``bits`` = bit-packed presence (torch int32 on the GPU), ``edges`` = (E, 2) family
        pairs; an edge has one gene pair in every genome that holds both families."""
        import torch
        device = bits.device if device is None else device
        F = int(bits.shape[0])
        E = len(edges)
        src = np.concatenate([edges[:, 0], edges[:, 1]]).astype(np.int64)
        dst = np.concatenate([edges[:, 1], edges[:, 0]]).astype(np.int64)
        eid = np.concatenate([np.arange(E), np.arange(E)]).astype(np.int64)
        order = np.argsort(src, kind="stable")
        adj_ptr = np.zeros(F + 1, np.int64)
        adj_ptr[1:] = np.cumsum(np.bincount(src, minlength=F))
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.int32)).to(device)
        return DevicePangenome(bits, G, t(adj_ptr), t(dst[order]), t(eid[order]), E, cov=None, edge_fam=t(edges), max_pairs=G)

    # ------------------------------------------------------------------------------------------------------ export
    def _c_pan(self) -> _Pan:
        p = engine._ptr
        return _Pan(self.F, self.G, self.E, self.A, self.Wg, p(self.presence), p(self.adj_ptr), p(self.adj_other), p(self.adj_edge),
                    p(self.cov[0]) if self.cov else 0, p(self.cov[1]) if self.cov else 0, p(self.cov[2]) if self.cov else 0,
                    p(self.edge_fam) if self.edge_fam is not None else 0)

    def weight_lut(self, n_sel: int):
        lut = self._luts.get(n_sel)
        if lut is None:
            cap = self.max_pairs if self.cov is not None else n_sel      # no pair count of a subset can exceed it
            lut = self.torch.from_numpy(python_round4_lut(n_sel, cap)).to(self.device)
            self._luts[n_sel] = lut
        return lut

    def columns(self, organisms) -> List[int]:
        """Genome columns of a collection of organism objects (or of column indices), ascending."""
        if self.org_index is None:
            return sorted(int(o) for o in organisms)
        return sorted(self.org_index[o] if not isinstance(o, (int, np.integer)) else int(o) for o in organisms)

    def export(self, samples: Sequence[Sequence[int]], sm_degree: float = 10, want_levels: bool = True) -> List[DeviceInstance]:
        """Instances of a batch of genome subsets (lists of ascending column indices, all of the same length; ``None`` = all
        genomes in column order) in one call of the CUDA exporter."""
        torch = self.torch
        n = len(samples)
        if n == 0:
            return []
        lens = {self.G if s is None else len(s) for s in samples}
        if len(lens) != 1:
            raise engine.NemB200Error("all genome subsets of one export call must have the same size")
        n_sel = lens.pop()
        identity = [s is None for s in samples]
        W = self.Wg if any(identity) else row_stride_words(n_sel)
        if any(identity) and not all(identity):
            raise engine.NemB200Error("the all-genomes instance is exported on its own")
        dev = self.device
        i32 = dict(dtype=torch.int32, device=dev)
        sel_t = None
        if not identity[0]:
            sel_t = torch.from_numpy(np.ascontiguousarray(np.asarray(samples, dtype=np.int32).reshape(n, n_sel))).to(dev)
        F, A = self.F, max(self.A, 1)
        bits = torch.empty((n, F, W), **i32)
        fam_rows = torch.empty((n, F), **i32)
        rowptr = torch.empty((n, F + 1), **i32)
        col = torch.empty((n, A), **i32)
        w = torch.empty((n, A), dtype=torch.float32, device=dev)
        lptr = torch.empty((n, F + 1), **i32) if want_levels else None
        lrows = torch.empty((n, F), **i32) if want_levels else None
        lut = self.weight_lut(n_sel)
        arr = (_Sample * n)()
        p = engine._ptr
        for i in range(n):
            arr[i] = _Sample(p(sel_t[i]) if sel_t is not None else 0, n_sel, W, p(bits[i]), p(fam_rows[i]), p(rowptr[i]), p(col[i]),
                             p(w[i]), p(lptr[i]) if want_levels else 0, p(lrows[i]) if want_levels else 0, 0, 0, 1, 0.0)
        pan = self._c_pan()
        with torch.cuda.device(dev):
            rc = engine.lib().nemb200_export_samples(ctypes.byref(pan), n, ctypes.cast(arr, _vp), float(sm_degree), p(lut),
                                                     int(lut.numel()), int(want_levels), engine._stream())
        engine._check(rc, "nemb200_export_samples")
        out = []
        for i in range(n):
            N, nnz, nl = int(arr[i].N), int(arr[i].nnz), int(arr[i].n_levels)
            di = engine.DeviceInput(bits[i, :N], n_sel)
            if nnz > 0:
                di.rowptr, di.col, di.w = rowptr[i, :N + 1], col[i, :nnz], w[i, :nnz]
                if want_levels:
                    di.sched_ptr, di.sched_rows, di.n_levels = lptr[i, :nl + 1], lrows[i, :N], nl
            out.append(DeviceInstance(di, fam_rows[i, :N], N, n_sel, float(arr[i].edges_weight)))
        return out

    def subset_counts(self, columns: Sequence[int]):
        """Per family the number of the given genomes that hold it (rarefaction.py:657-684 with gmpy2 bitarrays there)."""
        torch = self.torch
        sel = torch.from_numpy(np.ascontiguousarray(np.asarray(sorted(columns), dtype=np.int32))).to(self.device)
        mask = torch.empty(self.Wg, dtype=torch.int32, device=self.device)
        counts = torch.empty(self.F, dtype=torch.int32, device=self.device)
        pan = self._c_pan()
        with torch.cuda.device(self.device):
            rc = engine.lib().nemb200_subset_counts(ctypes.byref(pan), engine._ptr(sel), int(sel.numel()), engine._ptr(mask),
                                                    engine._ptr(counts), engine._stream())
        engine._check(rc, "nemb200_subset_counts")
        return counts


class ChunkVotes:
    """``validate_family`` (partition.py:784-802) on the device: votes[F][4] (P, S, C, U), validated flags and their count."""

    def __init__(self, F: int, n_org: int, chunk_size: int, device):
        import torch
        self.torch, self.F, self.n_org, self.ratio, self.device = torch, F, n_org, n_org / chunk_size, device
        self.votes = torch.zeros((F, 4), dtype=torch.int32, device=device)
        self.validated = torch.zeros(F, dtype=torch.uint8, device=device)
        self.n_validated = torch.zeros(1, dtype=torch.int32, device=device)

    def codes(self, T, K: int, fam_rows, out_row):
        """Partition letter of every family of one finished chunk into ``out_row`` (int8 [F], pre-filled with -1)."""
        with self.torch.cuda.device(self.device):
            rc = engine.lib().nemb200_chunk_codes(engine._ptr(T), int(T.shape[0]), K, engine._ptr(fam_rows), engine._ptr(out_row),
                                                  engine._stream())
        engine._check(rc, "nemb200_chunk_codes")

    def apply(self, codes):
        """Votes of the chunks of one round, in sample order (``codes``: int8 [n_chunks, F])."""
        if codes.shape[0] == 0:
            return
        codes = codes.contiguous()
        with self.torch.cuda.device(self.device):
            rc = engine.lib().nemb200_apply_votes(engine._ptr(codes), int(codes.shape[0]), self.F, engine._ptr(self.votes),
                                                  engine._ptr(self.validated), engine._ptr(self.n_validated), self.ratio, self.n_org,
                                                  engine._stream())
        engine._check(rc, "nemb200_apply_votes")

    def count_validated(self) -> int:
        return int(self.n_validated.item())

    def letters(self) -> np.ndarray:
        """Final partition letter per family: argmax of the votes, ties to the first of P, S, C, U like ``max(dict)``."""
        best = self.votes.argmax(1).cpu().numpy()
        return np.array(["P", "S", "C", "U"])[best]
