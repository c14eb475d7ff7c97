"""In-memory replacement for the temp-file export of ``write_nem_input_files``
(/root/reference/ppanggolin/nem/partition.py:344-436).

The reference writes ``nem_file.{str,dat,nei,index}`` as text and the C program parses them back
(NEM/nem_exe.c:839-903 matrix, :1347-1483 neighbours).  Here the same information goes straight into the arrays the
kernels consume: a bit-packed presence matrix and CSR neighbour lists holding exactly what the C reader would keep
(weights after ``round(.,4)`` and the float32 cast; zero weights dropped the way nem_exe.c:1413-1451 drops them).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from ..synthetic import row_stride_words


@dataclass
class NemInput:
    """One exported NEM instance (the content of one reference temp directory)."""
    bits: np.ndarray                 # (N, Ws) uint32
    D: int
    rowptr: np.ndarray               # (N+1,) int32
    col: np.ndarray                  # (nnz,) int32
    w: np.ndarray                    # (nnz,) float32
    fam_names: List[str]             # row -> family name (.index file)
    org_names: List[str]             # column -> genome name (column_org_file)
    edges_weight: float              # total_edges_weight / 2
    nei_text: Optional[list] = None  # per-row (indices, weight strings) as written to .nei, kept for keep_tmp_files
    device: Optional[object] = None  # engine.DeviceInput once uploaded
    bits_tensor: Optional[object] = None  # (N, Ws) int32 torch tensor when the instance was derived on the device
    fam_rows: Optional[object] = None     # torch int64 (N,): global family index of each row (array exporter only)

    @property
    def N(self) -> int:
        return len(self.fam_names)

    def host_bits(self) -> np.ndarray:
        if self.bits is None:
            self.bits = self.bits_tensor.cpu().numpy().view(np.uint32)
        return self.bits


_REGISTRY: Dict[str, NemInput] = {}


def register(tmpdir, inp: NemInput) -> None:
    _REGISTRY[str(Path(tmpdir))] = inp


def lookup(tmpdir) -> Optional[NemInput]:
    return _REGISTRY.get(str(Path(tmpdir)))


def release(prefix) -> None:
    """Drop every registered instance under ``prefix`` (the temp dir owned by ``partition()``)."""
    p = str(Path(prefix))
    for key in [k for k in _REGISTRY if k == p or k.startswith(p + "/")]:
        del _REGISTRY[key]


def c_reader_neighbours(idx: List[int], wtext: List[str]):
    """What NEM/nem_exe.c:1395-1451 keeps from one .nei row: indices are stored first, then weights are read and the
    zero ones skipped, so ``nv`` = number of non-zero weights and the kept pairs are (idx[t], t-th non-zero weight)."""
    wv = [np.float32(float(t)) for t in wtext]
    kept = [x for x in wv if x != 0.0]
    return idx[: len(kept)], kept


def export_from_pangenome(pan, organisms: set, sm_degree: float = 10, keep_text: bool = False) -> NemInput:
    """Same traversal as partition.py:373-436, producing arrays instead of files.

    Duck-types the reference data model: ``pan.gene_families``, ``fam.organisms``, ``fam.name``, ``fam.edges``,
    ``edge.source/target``, ``edge.get_organisms_dict()``, ``org.name``.
    """
    org_list = list(organisms)                      # column order = iteration order of the set (partition.py:377-379)
    index_org = {org: i for i, org in enumerate(org_list)}
    D = len(org_list)
    Ws = row_stride_words(D)
    index_fam = {}
    rows_cols = []
    for fam in pan.gene_families:                    # partition.py:381-391
        fam_organisms = set(fam.organisms)
        if not organisms.isdisjoint(fam_organisms):
            cols = np.fromiter((index_org[o] for o in (fam_organisms & organisms)), dtype=np.int64)
            rows_cols.append(cols)
            index_fam[fam] = len(index_fam)
    N = len(index_fam)
    bits = np.zeros((N, Ws), np.uint32)
    for r, cols in enumerate(rows_cols):
        np.bitwise_or.at(bits[r], cols >> 5, (np.uint32(1) << (cols & 31).astype(np.uint32)))

    total_edges_weight = 0.0
    rowptr = np.zeros(N + 1, np.int64)
    cols_out, w_out = [], []
    nei_text = [] if keep_text else None
    for fam, r in index_fam.items():                 # partition.py:393-433
        row_fam, row_w = [], []
        sum_dist_score = 0.0
        for edge in fam.edges:
            coverage = sum(len(gene_list) for org, gene_list in edge.get_organisms_dict().items() if org in organisms)
            if coverage == 0:
                continue
            distance_score = coverage / D
            sum_dist_score += distance_score
            row_fam.append(index_fam[edge.target if fam == edge.source else edge.source])
            row_w.append(str(round(distance_score, 4)))
        n = len(row_fam)
        if n > 0 and float(n) < sm_degree:
            total_edges_weight += sum_dist_score
            idx, kept = c_reader_neighbours(row_fam, row_w)
            cols_out.extend(idx); w_out.extend(kept)
            rowptr[r + 1] = len(idx)
            if keep_text:
                nei_text.append((row_fam, row_w))
        elif keep_text:
            nei_text.append(None)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return NemInput(bits, D, rowptr, np.asarray(cols_out, np.int32), np.asarray(w_out, np.float32),
                    [fam.name for fam in index_fam], [org.name for org in org_list], total_edges_weight / 2, nei_text)


def from_arrays(X_or_bits, D, rowptr, col, w, edges_weight, fam_names=None, org_names=None) -> NemInput:
    """Wrap already-exported arrays (synthetic generators, tests)."""
    from ..synthetic import pack_rows
    a = np.asarray(X_or_bits)
    bits = a.view(np.uint32) if a.dtype in (np.uint32, np.int32) else pack_rows(a)
    N = bits.shape[0]
    return NemInput(np.ascontiguousarray(bits), D, np.asarray(rowptr, np.int32), np.asarray(col, np.int32),
                    np.asarray(w, np.float32), fam_names or [f"fam{i}" for i in range(N)],
                    org_names or [f"org{j}" for j in range(D)], float(edges_weight))


def write_text_files(tmpdir: Path, inp: NemInput) -> None:
    """The reference's files (section 3.6 of SURVEY.md) for ``--keep_tmp_files`` and for diffing against the oracle."""
    from ..synthetic import unpack_rows
    tmpdir = Path(tmpdir)
    tmpdir.mkdir(parents=True, exist_ok=True)
    X = unpack_rows(inp.host_bits(), inp.D)
    with open(tmpdir / "column_org_file", "w") as f:
        f.write(" ".join(f'"{n}"' for n in inp.org_names) + "\n")
    with open(tmpdir / "nem_file.str", "w") as f:
        f.write(f"S\t{inp.N}\t{inp.D}\n")
    with open(tmpdir / "nem_file.dat", "w") as f:
        for row in X:
            f.write("\t".join("1" if v else "0" for v in row) + "\n")
    with open(tmpdir / "nem_file.index", "w") as f:
        for i, name in enumerate(inp.fam_names):
            f.write(f"{i + 1}\t{name}\n")
    with open(tmpdir / "nem_file.nei", "w") as f:
        f.write("1\n")
        for i in range(inp.N):
            ent = inp.nei_text[i] if inp.nei_text is not None else None
            if ent is None and inp.nei_text is None:
                a, b = int(inp.rowptr[i]), int(inp.rowptr[i + 1])
                if b > a:
                    ent = ([int(c) for c in inp.col[a:b]], [repr(round(float(v), 4)) for v in inp.w[a:b]])
            if ent:
                idx, wt = ent
                f.write("\t".join([str(i + 1), str(len(idx))] + [str(j + 1) for j in idx] + list(wt)) + "\n")
            else:
                f.write(f"{i + 1}\t0\n")
