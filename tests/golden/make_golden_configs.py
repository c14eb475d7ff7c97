#!/usr/bin/env python3
"""Golden vectors at the shapes BASELINE.json names (configs[1..3] and a wide-D instance), generated from the UNMODIFIED
reference C (oracle/_ref/libnem_ref.so through its file interface) where the reference is numerically valid (D <= 1000),
and from the fp64 restatement (oracle/nem_oracle.c, mode FP64) for the wide matrix, where the reference underflows
(SURVEY.md section 0-5).  Run in the build container (needs /root/reference); takes ~10 minutes on 8 cores.

    python tests/golden/make_golden_configs.py [cfg2] [cfg3] [cfg4] [wide]

The instances themselves are NOT stored: tests regenerate them from the seed (numpy generators, deterministic) and
check the digest recorded here before comparing results.
"""
import hashlib
import json
import shutil
import sys
import tempfile
import time
from multiprocessing import get_context
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refnem  # noqa: E402
from ppanggolin_b200.synthetic import synth_pangenome, synth_pangenome_blocked  # noqa: E402

OUT = Path(__file__).resolve().parent


def digest(sp) -> str:
    h = hashlib.sha256()
    for a in (np.ascontiguousarray(sp.bits), sp.rowptr, sp.col, sp.w):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:32]


write_files_fast = refnem.write_nem_files_fast


def _ref_one(args):
    d, K, beta, fd, itermax, D = args
    refnem.write_init_file(Path(d), K, D)
    t = time.time()
    rc = refnem.ref_nem_call(Path(d), K, float(np.float32(beta)), fd, itermax)
    dt = time.time() - t
    out = refnem.parse_ref_outputs(Path(d), K, D)
    return K, rc, dt, out


def mu_codes(mu):
    return np.where(mu == 1.0, 1, np.where(mu == 0.0, 0, 2)).astype(np.uint8)


def pack_result(out, with_T3=True):
    lab, _, ent = refnem.oracle_labels(out["T3"].astype(np.float32))
    d = dict(crit=np.array(out["crit"]), n_iter=out["n_iter"], converged=int(bool(out["converged"])), labels=lab.astype(np.int8),
             entropy=ent, mu=np.packbits(mu_codes(out["mu"]) == 1, axis=1), mu_half=int((mu_codes(out["mu"]) == 2).sum()),
             pk=out["pk"], eps=out["eps"][:, :8])
    if with_T3:
        d["T3"] = np.rint(out["T3"] * 1000).astype(np.uint16)
    return d


def cfg2():
    """BASELINE configs[1]: 1 000 genomes x 20 000 families, K = 3, beta 2.5, sk_, one run."""
    sp = synth_pangenome(20000, 1000, seed=42)
    beta = sp.beta_eff(2.5)
    td = Path(tempfile.mkdtemp())
    write_files_fast(td, sp.X, sp.rowptr, sp.col, sp.w)
    K, rc, dt, out = _ref_one((td, 3, beta, False, 100, sp.D))
    shutil.rmtree(td, ignore_errors=True)
    np.savez_compressed(OUT / "cfg2_ref.npz", digest=digest(sp), N=sp.N, D=sp.D, K=3, beta=np.float32(beta), rc=rc, ref_seconds=dt,
                        **pack_result(out))
    print("cfg2", rc, out["crit"], out["n_iter"], f"{dt:.1f}s")


def cfg3():
    """BASELINE configs[2], the K-selection part: 50 000 families x 500 sampled genomes, beta 0, 10 iterations, K = 2..20."""
    sp = synth_pangenome(50000, 500, seed=43)
    td = Path(tempfile.mkdtemp())
    write_files_fast(td, sp.X, sp.rowptr, sp.col, sp.w)
    ks = list(range(2, 21))
    with get_context("fork").Pool(8) as pool:
        res = pool.map(_ref_one, [(td, k, 0.0, False, 10, sp.D) for k in sorted(ks, reverse=True)])
    shutil.rmtree(td, ignore_errors=True)
    d = dict(digest=digest(sp), N=sp.N, D=sp.D, ks=np.array(ks))
    summary = {}
    for K, rc, dt, out in res:
        if out is None:
            summary[K] = dict(rc=rc, ok=0, seconds=dt)
            continue
        p = pack_result(out, with_T3=False)
        summary[K] = dict(rc=rc, ok=1, seconds=dt, M=out["crit"][3], Dcrit=out["crit"][1], n_iter=out["n_iter"], entropy=p["entropy"])
        if K in (2, 3, 8, 12, 20):
            d[f"labels_K{K}"] = p["labels"]
            d[f"pk_K{K}"] = p["pk"]
        print("cfg3 K", K, rc, out["crit"][3], out["n_iter"], f"{dt:.1f}s", flush=True)
    d["summary"] = json.dumps(summary)
    np.savez_compressed(OUT / "cfg3_ref.npz", **d)


def cfg4():
    """BASELINE configs[3], one chunk instance: 100 000 families x 500 genomes, K = 3, beta 2.5, free dispersion."""
    sp = synth_pangenome(100000, 500, seed=44)
    beta = sp.beta_eff(2.5)
    td = Path(tempfile.mkdtemp())
    write_files_fast(td, sp.X, sp.rowptr, sp.col, sp.w)
    K, rc, dt, out = _ref_one((td, 3, beta, True, 100, sp.D))
    shutil.rmtree(td, ignore_errors=True)
    np.savez_compressed(OUT / "cfg4_ref.npz", digest=digest(sp), N=sp.N, D=sp.D, K=3, beta=np.float32(beta), rc=rc, ref_seconds=dt,
                        **pack_result(out))
    print("cfg4", rc, out["crit"], out["n_iter"], f"{dt:.1f}s")


def wide():
    """A wide matrix (50 000 genomes x 20 000 families, un-chunked): the reference underflows here, the fp64 restatement
    (oracle/nem_oracle.c, FP64 mode, validated against the reference where both are valid) is the checker."""
    sp = synth_pangenome_blocked(20000, 50000, seed=45, keep_X=True)
    beta = sp.beta_eff(2.5)
    t = time.time()
    o = refnem.oracle_nem_run(sp.X, sp.rowptr, sp.col, sp.w, 3, beta, False, 100, mode=refnem.FP64)
    dt = time.time() - t
    lab, T3, ent = refnem.oracle_labels(o["T"])
    np.savez_compressed(OUT / "wide_fp64.npz", digest=digest(sp), N=sp.N, D=sp.D, K=3, beta=np.float32(beta), status=o["status"],
                        crit=np.array(o["crit"]), n_iter=o["n_iter"], converged=int(o["converged"]), labels=lab.astype(np.int8),
                        entropy=ent, T3=np.rint(T3.astype(np.float64) * 1000).astype(np.uint16), pk=o["pk"],
                        mu=np.packbits(o["mu"] == 1.0, axis=1), eps=o["eps"][:, :8], oracle_seconds=dt)
    print("wide", o["status"], o["crit"], o["n_iter"], f"{dt:.1f}s")


if __name__ == "__main__":
    refnem.build()
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg4", "wide"]
    for name in which:
        globals()[name]()
