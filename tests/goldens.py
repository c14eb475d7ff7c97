"""Loader for tests/golden/*.npz (generated from the unmodified reference by tests/golden/make_golden.py)."""
import json
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"
NEM_CASES = sorted(p.stem[4:] for p in GOLD.glob("nem_*.npz"))
PARTITION_CASES = sorted(p.stem[10:] for p in GOLD.glob("partition_*.npz") if p.stem != "partition_gbff")


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
    return dict(n_fam=int(z["n_fam"]), layouts=layouts, fam_names=[str(s) for s in z["fam_names"]],
                labels=[str(s) for s in z["labels"]], kwargs=json.loads(str(z["kwargs"])), params=json.loads(str(z["params"])))


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
