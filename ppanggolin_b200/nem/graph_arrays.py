"""The neighbours graph of a pangenome built on arrays.

Restates ``compute_neighbors_graph`` (/root/reference/ppanggolin/graph/makeGraph.py:87-138) for genomes given as arrays of
family ids in gene order instead of ``Organism`` / ``Contig`` / ``Gene`` objects: consecutive genes of a contig are linked
(:114-123; a pair of the same family is skipped when either gene is a fragment), a circular contig is closed (:130-132),
families with ``remove_copy_number`` or more copies in one genome are left out (:66-75, :103-104), and every gene pair is
recorded under its genome (``Pangenome.add_edge`` -> ``Edge.add_genes``, edge.py:98-120) -- the per-(edge, genome) pair
counts that ``write_nem_input_files`` turns into contiguity weights (partition.py:399-413).

The result is a :class:`export_arrays.PangenomeArrays` (presence matrix, adjacency in ``fam.edges`` order = order in which a
family's edges were first created, (edge, genome, pairs) table), i.e. exactly what one traversal of the reference's objects
gives; ``DevicePangenome.from_arrays`` puts it into HBM for the CUDA exporter.  Host-side numpy: the graph is built once.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def build_from_contigs(genomes: Sequence[Sequence[np.ndarray]], n_fam: int, fragments: Optional[Sequence[Sequence[np.ndarray]]] = None,
                       circular=True, remove_copy_number: int = 0, org_names: Optional[List[str]] = None, device="cpu"):
    """``genomes[g]`` = list of contigs, each an int array of family ids in gene order; ``fragments[g][c]`` = bool array (gene
    is a fragment), ``circular`` = bool or per-genome list of per-contig bools.  Returns ``export_arrays.PangenomeArrays``."""
    import torch
    from .export_arrays import PangenomeArrays
    G = len(genomes)
    # families that occur, in ascending id order (the order a pangenome built from these contigs iterates them in)
    all_ids = np.concatenate([np.asarray(c, np.int64) for g in genomes for c in g]) if G else np.zeros(0, np.int64)
    used = np.unique(all_ids)
    fam_of = np.full(n_fam, -1, np.int64)
    fam_of[used] = np.arange(len(used))
    F = len(used)
    presence = np.zeros((F, G), bool)
    removed = np.zeros(F, bool)
    for g, contigs in enumerate(genomes):
        ids = fam_of[np.concatenate([np.asarray(c, np.int64) for c in contigs])] if len(contigs) else np.zeros(0, np.int64)
        presence[ids, g] = True
        if remove_copy_number > 0 and len(ids):
            removed |= np.bincount(ids, minlength=F) >= remove_copy_number          # makeGraph.py:66-75
    pa, pb, pg = [], [], []                                                         # gene pairs in creation order
    for g, contigs in enumerate(genomes):
        for ci, c in enumerate(contigs):
            fam = fam_of[np.asarray(c, np.int64)]
            n = len(fam)
            if n == 0:
                continue
            frag = np.asarray(fragments[g][ci], bool) if fragments is not None else np.zeros(n, bool)
            circ = circular if isinstance(circular, bool) else bool(circular[g][ci])
            keep = ~removed[fam]
            kf, kfr = fam[keep], frag[keep]
            if len(kf) >= 2:                                                        # makeGraph.py:114-123
                a, b = kf[1:], kf[:-1]
                ok = ~((a == b) & (kfr[1:] | kfr[:-1]))
                pa.append(a[ok]); pb.append(b[ok]); pg.append(np.full(int(ok.sum()), g, np.int64))
            if len(kf) >= 1 and circ:                                               # makeGraph.py:130-132: add_edge(contig[0], prev)
                pa.append(np.array([fam[0]])); pb.append(np.array([kf[-1]])); pg.append(np.array([g], np.int64))
    if pa:
        pa, pb, pg = np.concatenate(pa), np.concatenate(pb), np.concatenate(pg)
    else:
        pa = pb = pg = np.zeros(0, np.int64)
    lo, hi = np.minimum(pa, pb), np.maximum(pa, pb)
    key = lo * max(F, 1) + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                                        # edge ids in creation order
    rank = np.empty(len(uniq), np.int64); rank[order] = np.arange(len(uniq))
    eid = rank[inv]
    E = len(uniq)
    e_lo, e_hi = (uniq[order] // max(F, 1)), (uniq[order] % max(F, 1))
    # adjacency: every edge enters the edge dict of both its families when it is created (a self-loop once)
    owner = np.concatenate([e_lo, e_hi[e_lo != e_hi]])
    other = np.concatenate([e_hi, e_lo[e_lo != e_hi]])
    edge = np.concatenate([np.arange(E), np.arange(E)[e_lo != e_hi]])
    o = np.lexsort((edge, owner))                                                   # per family, edges in creation order
    owner, other, edge = owner[o], other[o], edge[o]
    adj_ptr = np.zeros(F + 1, np.int64)
    adj_ptr[1:] = np.cumsum(np.bincount(owner, minlength=F))
    # gene pairs per (edge, genome)
    ck = eid * max(G, 1) + pg
    cu, cc = np.unique(ck, return_counts=True)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x, np.int64)).to(device)
    names = org_names if org_names is not None else [f"org{g}" for g in range(G)]
    return PangenomeArrays([f"fam{int(i)}" for i in used], names, {g: g for g in range(G)}, torch.from_numpy(presence).to(device),
                           t(adj_ptr), t(owner), t(edge), t(other), t(cu // max(G, 1)), t(cu % max(G, 1)), t(cc), E, device)
