"""Array form of a whole pangenome, from which every NEM instance (all genomes, the K-selection sample, each chunk of
the chunked mode) is derived with tensor operations instead of a new traversal of the Python objects.

The reference walks every family and every edge again for every sample (``write_nem_input_files``,
/root/reference/ppanggolin/nem/partition.py:373-436; ~1.2 us per matrix cell), which is what a
chunked ``partition()`` spends its time on once the EM itself is fast.  Here the objects are read ONCE
(:meth:`PangenomeArrays.from_pangenome`) into

* ``presence``  bool [families, genomes]
* the edge table: per edge its two families, and a COO list (edge, genome, number of gene pairs)
* the adjacency in ``fam.edges`` order: (owner family, edge, other family)

and :meth:`PangenomeArrays.instance` produces, for a genome subset, exactly the arrays ``export_from_pangenome`` would
(same row order, same neighbour order, same ``round(coverage / n, 4)`` weights, same ``sm_degree`` cut, same zero-weight
behaviour of the C reader, same ``total_edges_weight / 2``) -- checked against it in tests/test_export_arrays.py.
Everything is torch, so the same code runs on the GPU in production and on CPU tensors in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from ..synthetic import row_stride_words
from .export import NemInput


@dataclass
class PangenomeArrays:
    fam_names: List[str]
    org_names: List[str]
    org_index: dict                 # organism object -> column
    presence: object                # torch bool [F, G]
    adj_ptr: object                 # int64 [F+1]  adjacency rows (fam.edges order)
    adj_fam: object                 # int64 [A]    owner family of each adjacency entry
    adj_edge: object                # int64 [A]
    adj_other: object               # int64 [A]
    cov_edge: object                # int64 [M]    COO over (edge, genome)
    cov_org: object                 # int64 [M]
    cov_cnt: object                 # int64 [M]    gene pairs of that edge in that genome
    n_edges: int
    device: object

    @staticmethod
    def from_pangenome(pan, device="cpu") -> "PangenomeArrays":
        import torch
        orgs = list(pan.organisms)
        org_index = {o: i for i, o in enumerate(orgs)}
        fams = list(pan.gene_families)
        fam_index = {f: i for i, f in enumerate(fams)}
        F, G = len(fams), len(orgs)
        cols, counts = [], []
        for fam in fams:
            n0 = len(cols)
            cols.extend([org_index[o] for o in fam.organisms])
            counts.append(len(cols) - n0)
        presence = torch.zeros((F, G), dtype=torch.bool)
        if cols:   # numpy builds the index arrays (torch.tensor on a long Python list is several times slower)
            rows = np.repeat(np.arange(F, dtype=np.int64), np.asarray(counts, dtype=np.int64))
            presence[torch.from_numpy(rows), torch.from_numpy(np.asarray(cols, dtype=np.int64))] = True
        edge_index = {}
        adj_ptr = [0]
        adj_fam, adj_edge, adj_other = [], [], []
        cov_edge, cov_org, cov_cnt = [], [], []
        for i, fam in enumerate(fams):
            for edge in fam.edges:
                eid = edge_index.get(id(edge))
                if eid is None:
                    eid = len(edge_index)
                    edge_index[id(edge)] = eid
                    for org, gene_list in edge.get_organisms_dict().items():
                        cov_edge.append(eid); cov_org.append(org_index[org]); cov_cnt.append(len(gene_list))
                other = edge.target if fam == edge.source else edge.source
                adj_fam.append(i); adj_edge.append(eid); adj_other.append(fam_index[other])
            adj_ptr.append(len(adj_fam))
        t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int64)).to(device)
        return PangenomeArrays([f.name for f in fams], [o.name for o in orgs], org_index, presence.to(device), t(adj_ptr),
                               t(adj_fam), t(adj_edge), t(adj_other), t(cov_edge), t(cov_org), t(cov_cnt), len(edge_index),
                               device)

    # ------------------------------------------------------------------------------------------------------------
    def instance(self, organisms, sm_degree: float = 10) -> NemInput:
        """The NEM instance of a genome subset (iterable of organism objects or of column indices)."""
        import torch
        dev = self.device
        sel = sorted(self.org_index[o] if not isinstance(o, (int, np.integer)) else int(o) for o in organisms)
        S = torch.tensor(sel, dtype=torch.long, device=dev)
        Ds = len(sel)
        F = self.presence.shape[0]
        # ---- matrix rows: families present in at least one selected genome, in family order (partition.py:381-391)
        sub = self.presence.index_select(1, S)
        present = sub.any(1)
        fam_rows = present.nonzero().squeeze(1)
        N = int(fam_rows.numel())
        rowmap = torch.cumsum(present.to(torch.long), 0) - 1
        W = row_stride_words(Ds)
        shifts = torch.arange(32, device=dev, dtype=torch.int64)
        bits = torch.empty((N, W), dtype=torch.int32, device=dev)
        step = max(1, (1 << 24) // max(W * 32, 1))          # rows per block: the int64 expansion stays around 128 MB
        for r0 in range(0, N, step):
            rows = fam_rows[r0:r0 + step]
            x = torch.zeros((rows.numel(), W * 32), dtype=torch.bool, device=dev)
            x[:, :Ds] = sub.index_select(0, rows)
            words = (x.view(rows.numel(), W, 32).to(torch.int64) << shifts).sum(-1)
            bits[r0:r0 + step] = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
        # ---- edge coverage in the subset (partition.py:399-407)
        mask = torch.zeros(self.presence.shape[1], dtype=torch.long, device=dev)
        mask[S] = 1
        cov = torch.zeros(max(self.n_edges, 1), dtype=torch.long, device=dev)
        if self.cov_edge.numel():
            cov.index_add_(0, self.cov_edge, self.cov_cnt * mask[self.cov_org])
        cov_a = cov[self.adj_edge] if self.adj_edge.numel() else torch.zeros(0, dtype=torch.long, device=dev)
        keep_a = cov_a > 0
        A = int(keep_a.numel())
        # ---- per-family neighbour count, sm_degree cut (partition.py:415)
        c_keep = torch.cumsum(keep_a.to(torch.long), 0)
        c0 = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), c_keep])
        n_nei = c0[self.adj_ptr[1:]] - c0[self.adj_ptr[:-1]]
        valid_fam = (n_nei > 0) & (n_nei.to(torch.float64) < float(sm_degree))
        ent = keep_a & valid_fam[self.adj_fam] if A else keep_a
        dist = cov_a.to(torch.float64) / Ds                                   # partition.py:408
        edges_weight = float(dist[ent].sum().item()) / 2 if A else 0.0        # partition.py:416,436
        # ---- weights: Python's round(coverage / n, 4) (partition.py:413), exact through a table over the distinct counts
        uniq = torch.unique(cov_a[ent]) if A else cov_a
        table = {int(c): float(np.float32(float(str(round(int(c) / Ds, 4))))) for c in uniq.tolist()}
        maxc = int(uniq.max().item()) if uniq.numel() else 0
        lut_host = np.zeros(maxc + 1, np.float32)
        for c, v in table.items():
            lut_host[c] = v
        lut = torch.from_numpy(lut_host).to(dev)
        wr = lut[cov_a.clamp(max=maxc)] if A else torch.zeros(0, dtype=torch.float32, device=dev)
        # ---- what the C reader keeps (nem_exe.c:1395-1451): the first nv indices, the non-zero weights in order
        nz = ent & (wr != 0)
        c_ent = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(ent.to(torch.long), 0)])
        c_nz = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(nz.to(torch.long), 0)])
        nv_fam = c_nz[self.adj_ptr[1:]] - c_nz[self.adj_ptr[:-1]]              # per family (all families)
        nv_rows = nv_fam[fam_rows]
        rowptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(nv_rows, 0)])
        nnz = int(rowptr[-1].item())
        col = torch.zeros(nnz, dtype=torch.int32, device=dev)
        w = torch.zeros(nnz, dtype=torch.float32, device=dev)
        if A and nnz:
            start = self.adj_ptr[:-1][self.adj_fam]                            # first adjacency slot of the owner family
            rank_ent = c_ent[:-1] - c_ent[start]                               # order among the family's kept entries
            rank_nz = c_nz[:-1] - c_nz[start]
            base = torch.zeros(F, dtype=torch.long, device=dev)
            base[fam_rows] = rowptr[:-1]
            take_i = ent & (rank_ent < nv_fam[self.adj_fam])
            col[(base[self.adj_fam] + rank_ent)[take_i]] = rowmap[self.adj_other][take_i].to(torch.int32)
            w[(base[self.adj_fam] + rank_nz)[nz]] = wr[nz]
        inp = NemInput(bits=None, D=Ds, rowptr=rowptr.to(torch.int32).cpu().numpy(), col=col.cpu().numpy(), w=w.cpu().numpy(),
                       fam_names=[self.fam_names[i] for i in fam_rows.tolist()], org_names=[self.org_names[i] for i in sel],
                       edges_weight=edges_weight)
        inp.bits_tensor = bits
        inp.fam_rows = fam_rows          # global family index of every instance row (vectorised chunk voting)
        return inp
