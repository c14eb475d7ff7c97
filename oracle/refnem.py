"""Test-infrastructure bindings for the NEM oracle.  NOT part of the product (ppanggolin_b200 never imports this).

Two checkers live here:

* ``ref_nem_run``   -- drives the UNMODIFIED reference C (``oracle/_ref/libnem_ref.so``, built by ``oracle/Makefile``
  from the sources under /root/reference) exactly the way ``ppanggolin/nem/partition.py:65-150`` does: text files in,
  ``nem()`` (NEM/nem_exe.h:23-38) through ctypes instead of the Cython shim (NEM/nem_stats.pyx:1-17), text files out.
* ``oracle_nem_run`` -- our C restatement (``oracle/libnem_oracle.so``; nem_oracle.c), in memory.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "libnem_ref.so"
ORACLE_SO = HERE / "libnem_oracle.so"


def build(ref: bool = True) -> None:
    """Compile the oracle (always) and the reference .so (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    if ref and Path("/root/reference/ppanggolin/nem/NEM/nem_exe.c").is_file():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True, stderr=subprocess.DEVNULL)


def have_ref() -> bool:
    return REF_SO.is_file()


# --------------------------------------------------------------------------- reference C through its file interface
_ref_lib = None


def _ref():
    global _ref_lib
    if _ref_lib is None:
        lib = ctypes.CDLL(str(REF_SO))
        lib.nem.restype = ctypes.c_int
        lib.nem.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_float, ctypes.c_char_p,
                            ctypes.c_float, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                            ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p,
                            ctypes.c_int]
        _ref_lib = lib
    return _ref_lib


def write_nem_files(dirpath: Path, X: np.ndarray, nei_lists, names=None) -> None:
    """Files of partition.py:344-436.  ``nei_lists[i]`` = (list of 0-based neighbour rows, list of weight STRINGS)
    or None / empty for a row written as ``idx\\t0``."""
    dirpath.mkdir(parents=True, exist_ok=True)
    N, D = X.shape
    with open(dirpath / "nem_file.str", "w") as f:
        f.write(f"S\t{N}\t{D}\n")
    with open(dirpath / "nem_file.dat", "w") as f:
        for row in X:
            f.write("\t".join("1" if v else "0" for v in row) + "\n")
    with open(dirpath / "nem_file.index", "w") as f:
        for i in range(N):
            f.write(f"{i + 1}\t{names[i] if names is not None else 'fam' + str(i)}\n")
    with open(dirpath / "nem_file.nei", "w") as f:
        f.write("1\n")
        for i in range(N):
            ent = nei_lists[i] if nei_lists is not None else None
            if ent and len(ent[0]) > 0:
                idx, wts = ent
                f.write("\t".join([str(i + 1), str(len(idx))] + [str(j + 1) for j in idx] + [str(w) for w in wts]) + "\n")
            else:
                f.write(f"{i + 1}\t0\n")


def write_nem_files_fast(dirpath: Path, X: np.ndarray, rowptr, col, w) -> None:
    """Same files as :func:`write_nem_files` from CSR arrays, vectorised for 10^7..10^8 cells (weights written the way
    partition.py:413 does: ``str(round(x, 4))``)."""
    dirpath.mkdir(parents=True, exist_ok=True)
    X = np.ascontiguousarray(X, dtype=np.uint8)
    N, D = X.shape
    (dirpath / "nem_file.str").write_text(f"S\t{N}\t{D}\n")
    buf = np.full((N, 2 * D), ord("\t"), np.uint8)
    buf[:, 0::2] = X + ord("0")
    buf[:, -1] = ord("\n")
    (dirpath / "nem_file.dat").write_bytes(buf.tobytes())
    with open(dirpath / "nem_file.index", "w") as f:
        f.write("".join(f"{i + 1}\tfam{i}\n" for i in range(N)))
    w_text = [repr(round(float(v), 4)) for v in np.asarray(w, np.float64)]
    with open(dirpath / "nem_file.nei", "w") as f:
        f.write("1\n")
        for i in range(N):
            a, b = int(rowptr[i]), int(rowptr[i + 1])
            if b > a:
                f.write("\t".join([str(i + 1), str(b - a)] + [str(int(j) + 1) for j in col[a:b]] + w_text[a:b]) + "\n")
            else:
                f.write(f"{i + 1}\t0\n")


def write_init_file(dirpath: Path, K: int, D: int) -> None:
    """partition.py:65-85, same formatting."""
    with open(dirpath / f"nem_file_init_{K}.m", "w") as m_file:
        m_file.write("1 ")
        base = format(1.0 / float(K), ".12g")
        m_file.write(" ".join([base] * (K - 1)) + " ")
        mu, eps = [], []
        step = 0.5 / (math.ceil(K / 2))
        pich = 0.1 if K == 2 else 0
        for k in range(1, K + 1):
            if k <= K / 2:
                mu += ["1"] * D
                eps += [str((step * k) - pich)] * D
            else:
                mu += ["0"] * D
                eps += [str((step * (K - k + 1)) - pich)] * D
        m_file.write(" ".join(mu) + " " + " ".join(eps))


def ref_nem_call(dirpath: Path, K: int, beta: float, free_dispersion: bool, itermax: int = 100, seed: int = 42) -> int:
    """The nem_stats.nem(...) call of partition.py:126-150."""
    p = dirpath.as_posix().encode("ascii")
    return _ref().nem(p + b"/nem_file", K, b"nem", beta, b"clas", 0.01, b"fuzzy", itermax, 1, b"bern", b"pk",
                      b"skd" if free_dispersion else b"sk_", 2, p + b"/nem_file_init_" + str(K).encode() + b".m",
                      p + b"/nem_file_" + str(K).encode(), seed)


def parse_ref_outputs(dirpath: Path, K: int, D: int):
    """Raw parse of .uf / .mf / .stderr (what partition.py:170-231 reads, kept numeric)."""
    uf = dirpath / f"nem_file_{K}.uf"
    if not uf.is_file():
        return None
    T3 = np.loadtxt(uf, dtype=np.float64, ndmin=2)
    lines = (dirpath / f"nem_file_{K}.mf").read_text().splitlines()
    crit = [float(v) for v in lines[2].split()[:5]]
    mu, pk, eps = [], [], []
    for line in lines[-K:]:
        v = line.split()
        mu.append([float(t) for t in v[0:D]])
        pk.append(float(v[D]))
        eps.append([float(t) for t in v[D + 1:]])
    n_iter, converged = None, None
    err = (dirpath / f"nem_file_{K}.stderr")
    if err.is_file():
        for line in err.read_text(errors="replace").splitlines():
            if "NEM converged after" in line:
                n_iter, converged = int(line.split()[3]), True
            elif "NEM did not converge after" in line:
                n_iter, converged = int(line.split()[5]), False
    return {"T3": T3, "crit": crit, "mu": np.array(mu), "pk": np.array(pk), "eps": np.array(eps),
            "n_iter": n_iter, "converged": converged}


def csr_to_nei_lists(N, rowptr, col, w_text):
    """CSR with weight STRINGS -> per-row lists for write_nem_files."""
    out = []
    for i in range(N):
        a, b = int(rowptr[i]), int(rowptr[i + 1])
        out.append(([int(c) for c in col[a:b]], list(w_text[a:b])) if b > a else None)
    return out


def ref_nem_run(X, rowptr, col, w, K, beta, free_dispersion=False, itermax=100, workdir=None, keep=False):
    """Reference run on arrays.  ``w`` are float32 weights that are exactly representable by their ``%.4f``-style
    short repr (the exporter only produces round(x, 4) values), written with Python ``str(float)`` as partition.py:413."""
    X = np.asarray(X, dtype=np.uint8)
    N, D = X.shape
    tmp = Path(tempfile.mkdtemp(dir=workdir)) if not keep else Path(workdir)
    w_text = [repr(round(float(v), 4)) for v in np.asarray(w, dtype=np.float64)] if len(w) else []
    write_nem_files(tmp, X, csr_to_nei_lists(N, rowptr, col, w_text))
    write_init_file(tmp, K, D)
    rc = ref_nem_call(tmp, K, float(np.float32(beta)), free_dispersion, itermax)
    out = parse_ref_outputs(tmp, K, D)
    if not keep:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return rc, out


# --------------------------------------------------------------------------- our C restatement, in memory
class _Problem(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("D", ctypes.c_int32), ("K", ctypes.c_int32),
                ("X", ctypes.c_void_p), ("rowptr", ctypes.c_void_p), ("col", ctypes.c_void_p), ("w", ctypes.c_void_p),
                ("beta", ctypes.c_float), ("free_disp", ctypes.c_int32), ("itermax", ctypes.c_int32),
                ("cv_th", ctypes.c_float), ("mode", ctypes.c_int32), ("jacobi", ctypes.c_int32)]


class _Result(ctypes.Structure):
    _fields_ = [("T", ctypes.c_void_p), ("mu", ctypes.c_void_p), ("eps", ctypes.c_void_p), ("pk", ctypes.c_void_p),
                ("crit", ctypes.c_double * 5), ("n_iter", ctypes.c_int32), ("converged", ctypes.c_int32),
                ("status", ctypes.c_int32)]


_oracle_lib = None


def _oracle():
    global _oracle_lib
    if _oracle_lib is None:
        if not ORACLE_SO.is_file():
            build(ref=False)
        lib = ctypes.CDLL(str(ORACLE_SO))
        lib.nem_oracle_run.restype = ctypes.c_int
        lib.nem_oracle_run.argtypes = [ctypes.POINTER(_Problem), ctypes.POINTER(_Result)]
        lib.nem_oracle_init_params.restype = None
        lib.nem_oracle_init_params.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.nem_oracle_labels.restype = ctypes.c_double
        lib.nem_oracle_labels.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _oracle_lib = lib
    return _oracle_lib


STRICT, FP64 = 0, 1


def oracle_init_params(K: int, D: int):
    pk = np.zeros(K, np.float32); mu = np.zeros((K, D), np.float32); eps = np.zeros((K, D), np.float32)
    _oracle().nem_oracle_init_params(K, D, pk.ctypes.data, mu.ctypes.data, eps.ctypes.data)
    return pk, mu, eps


def oracle_nem_run(X, rowptr, col, w, K, beta, free_dispersion=False, itermax=100, mode=STRICT, jacobi=False, cv_th=0.01):
    X = np.ascontiguousarray(X, dtype=np.uint8)
    N, D = X.shape
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    pk, mu, eps = oracle_init_params(K, D)
    T = np.zeros((N, K), np.float32)
    prob = _Problem(N, D, K, X.ctypes.data, rowptr.ctypes.data, col.ctypes.data, w.ctypes.data,
                    float(np.float32(beta)), int(bool(free_dispersion)), itermax, cv_th, mode, int(bool(jacobi)))
    res = _Result(T.ctypes.data, mu.ctypes.data, eps.ctypes.data, pk.ctypes.data)
    status = _oracle().nem_oracle_run(ctypes.byref(prob), ctypes.byref(res))
    return {"status": status, "T": T, "mu": mu, "eps": eps, "pk": pk, "crit": list(res.crit),
            "n_iter": res.n_iter, "converged": bool(res.converged)}


def oracle_labels(T):
    T = np.ascontiguousarray(T, dtype=np.float32)
    N, K = T.shape
    lab = np.zeros(N, np.int32); T3 = np.zeros((N, K), np.float32)
    ent = _oracle().nem_oracle_labels(N, K, T.ctypes.data, lab.ctypes.data, T3.ctypes.data)
    return lab, T3, ent


def fmt_g(x: float) -> float:
    """Value after a C ``%g`` print/parse round trip (6 significant digits), as partition.py:187 sees criteria."""
    return float("%g" % x)
