"""Loader for tests/golden/*.npz (generated from the unmodified reference by tests/golden/make_golden.py)."""
import json
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"
NEM_CASES = sorted(p.stem[4:] for p in GOLD.glob("nem_*.npz"))
PARTITION_CASES = sorted(p.stem[10:] for p in GOLD.glob("partition_*.npz")
                         if p.stem != "partition_gbff" and "replay" not in p.stem)
REPLAY_CASES = sorted(p.stem[10:] for p in GOLD.glob("partition_*replay*.npz"))


def load_nem(name):
    z = np.load(GOLD / f"nem_{name}.npz")
    g = {k: z[k] for k in z.files}
    for k in ("D", "K", "fd", "itermax", "rc", "ok"):
        g[k] = int(g[k])
    g["beta"] = float(g["beta"])
    if g["ok"]:
        g["n_iter"] = int(g["n_iter"]); g["converged"] = bool(g["converged"])
        g["T3"] = g["T3"].astype(np.float64) / 1000.0
    return g


def load_partition(name):
    z = np.load(GOLD / f"partition_{name}.npz")
    ptr = z["layout_ptr"]
    layouts = [z["layouts"][ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]
    out = dict(n_fam=int(z["n_fam"]), layouts=layouts, fam_names=[str(s) for s in z["fam_names"]],
               labels=[str(s) for s in z["labels"]], kwargs=json.loads(str(z["kwargs"])), params=json.loads(str(z["params"])))
    if "shuffles" in z.files:
        out["shuffles"] = json.loads(str(z["shuffles"]))
        out["n_samples"] = int(z["n_samples"])
    return out


def load_config(name):
    """tests/golden/{cfg2_ref,cfg3_ref,cfg4_ref,wide_fp64}.npz (tests/golden/make_golden_configs.py)."""
    z = np.load(GOLD / f"{name}.npz")
    g = {k: z[k] for k in z.files}
    for k in ("N", "D", "K", "n_iter", "converged", "rc", "status", "mu_half"):
        if k in g:
            g[k] = int(g[k])
    for k in ("beta", "entropy", "ref_seconds", "oracle_seconds"):
        if k in g:
            g[k] = float(g[k])
    g["digest"] = str(g["digest"])
    if "summary" in g:
        g["summary"] = {int(k): v for k, v in json.loads(str(g["summary"])).items()}
    return g


def instance_digest(sp) -> str:
    import hashlib
    h = hashlib.sha256()
    for a in (np.ascontiguousarray(sp.bits), sp.rowptr, sp.col, sp.w):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:32]


def g6(x):
    return float("%g" % x)


def load_partition_gbff():
    """BASELINE config 1 surrogate: reference genomes' gene order (several contigs per genome) + the reference's labels."""
    z = np.load(GOLD / "partition_gbff.npz")
    cp, gp = z["contig_ptr"], z["genome_ptr"]
    contigs = [z["families"][cp[i]:cp[i + 1]] for i in range(len(cp) - 1)]
    genomes = [contigs[gp[g]:gp[g + 1]] for g in range(len(gp) - 1)]
    return dict(n_fam=int(z["n_fam"]), genomes=genomes, fam_names=[str(s) for s in z["fam_names"]],
                labels=[str(s) for s in z["labels"]], kwargs=json.loads(str(z["kwargs"])), params=json.loads(str(z["params"])))
