#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container, where /root/reference exists):

* nem_*.npz       -- reference nem() (oracle/_ref/libnem_ref.so, file interface) on seeded instances: 3-decimal
                     posteriors, criteria, parameters, iteration counts, failure statuses.
* partition_*.npz -- reference ppanggolin/nem/partition.py::partition() on in-memory reference Pangenome objects
                     (import shim of SURVEY.md Appendix C): family -> label, chosen K, recorded parameters.

Usage:  python tests/golden/make_golden.py
"""
import json
import random
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refnem  # noqa: E402
from ppanggolin_b200.synthetic import pack_rows, synth_pangenome  # noqa: E402

OUT = Path(__file__).resolve().parent


def appendix_b():
    random.seed(1)
    N, D = 300, 40
    X = np.zeros((N, D), np.uint8)
    for i in range(N):
        r = random.random()
        p = 0.97 if r < 0.4 else (0.5 if r < 0.6 else 0.05)
        row = [1 if random.random() < p else 0 for _ in range(D)]
        if sum(row) == 0:
            row[0] = 1
        X[i] = row
    rowptr = [0]; col = []; w = []
    for i in range(N):
        for j in (i - 1, i + 1):
            if 0 <= j < N:
                col.append(j); w.append(0.5)
        rowptr.append(len(col))
    return X, np.array(rowptr, np.int32), np.array(col, np.int32), np.array(w, np.float32)


def save_nem_case(name, X, rowptr, col, w, K, beta, fd, itermax):
    rc, ref = refnem.ref_nem_run(X, rowptr, col, w, K, beta, fd, itermax)
    d = dict(bits=pack_rows(X), D=X.shape[1], rowptr=rowptr, col=col, w=w, K=K, beta=np.float32(beta), fd=int(fd),
             itermax=itermax, rc=rc, ok=int(ref is not None))
    if ref is not None:
        d.update(T3=np.rint(ref["T3"] * 1000).astype(np.uint16), crit=np.array(ref["crit"]), mu=ref["mu"].astype(np.float32),
                 pk=ref["pk"], eps=ref["eps"], n_iter=ref["n_iter"], converged=int(ref["converged"]))
    np.savez_compressed(OUT / f"nem_{name}.npz", **d)
    print(name, "rc", rc, "crit", ref["crit"] if ref else None, "iters", ref["n_iter"] if ref else None)


def nem_cases():
    X, rp, c, w = appendix_b()
    save_nem_case("appendixB", X, rp, c, w, 3, 2.5, False, 100)
    sp = synth_pangenome(1200, 150, seed=11)
    save_nem_case("sk_graph", sp.X, sp.rowptr, sp.col, sp.w, 3, sp.beta_eff(2.5), False, 100)
    save_nem_case("skd_graph", sp.X, sp.rowptr, sp.col, sp.w, 3, sp.beta_eff(2.6), True, 100)
    sp = synth_pangenome(1500, 120, seed=12)
    for K in (2, 3, 5, 8, 12, 20):
        save_nem_case(f"sweep_K{K}", sp.X, sp.rowptr, sp.col, sp.w, K, 0.0, False, 10)
    save_nem_case("sweep_fd_K7", sp.X, sp.rowptr, sp.col, sp.w, 7, 0.0, True, 10)
    sp = synth_pangenome(800, 1000, seed=13)
    save_nem_case("wide1000", sp.X, sp.rowptr, sp.col, sp.w, 3, sp.beta_eff(2.5), False, 100)
    sp = synth_pangenome(600, 64, seed=14, chain_order=True)
    save_nem_case("chain_order", sp.X, sp.rowptr, sp.col, sp.w, 4, sp.beta_eff(2.5), False, 100)
    # neighbour term with more classes than the 3-class runs above (9 classes: 16 lanes per row in the sweep kernel) ...
    sp = synth_pangenome(1400, 130, seed=15)
    save_nem_case("graph_K9", sp.X, sp.rowptr, sp.col, sp.w, 9, sp.beta_eff(2.5), False, 100)
    # ... and with neighbour lists of up to ~40 entries (sm_degree 60; the sweep keeps 8 inline per row)
    sp = synth_pangenome(1500, 120, seed=31, sm_degree=60)
    save_nem_case("highdeg", sp.X, sp.rowptr, sp.col, sp.w, 3, sp.beta_eff(2.5), False, 100)
    # a run the reference abandons (STS_W_EMPTYCLASS, no .uf): with 2000 columns the eps = 0.5 class underflows to an
    # exactly-zero posterior for every row, so its size is 0 at the first M-step
    X = np.zeros((40, 2000), np.uint8); X[:20] = 1
    none = np.zeros(41, np.int32)
    save_nem_case("emptyclass", X, none, np.zeros(0, np.int32), np.zeros(0, np.float32), 3, 0.0, False, 10)


# ------------------------------------------------------------------------------------------ reference partition()
def import_reference_partition():
    import importlib.abc
    import importlib.machinery
    import types
    from unittest.mock import MagicMock
    STUB_ROOTS = ("tables", "gmpy2", "plotly", "pyrodigal", "bokeh")

    class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in STUB_ROOTS:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)

        def create_module(self, spec):
            m = MagicMock(name=spec.name); m.__path__ = []; m.__spec__ = spec
            return m

        def exec_module(self, module):
            pass
    sys.meta_path.insert(0, _StubFinder())
    pkg = types.ModuleType("ppanggolin"); pkg.__path__ = ["/root/reference/ppanggolin"]
    sys.modules["ppanggolin"] = pkg
    # nem_stats: the Cython shim is a 1:1 forward to nem(); bind the same C symbol with ctypes
    ns = types.ModuleType("nem_stats")

    def nem(Fname, nk, algo, beta, convergence, convergence_th, format, it_max, dolog, model_family, proportion,
            dispersion, init_mode, init_file, out_file_prefix, seed):
        return refnem._ref().nem(Fname, nk, algo, beta, convergence, convergence_th, format, it_max, int(dolog),
                                 model_family, proportion, dispersion, init_mode, init_file, out_file_prefix, seed)
    ns.nem = nem
    sys.modules["nem_stats"] = ns
    import ppanggolin.nem.partition as ref_partition
    return ref_partition


def make_layouts(n_fam, n_org, seed):
    """Gene order of every genome: one circular contig over its present families (block-shuffled backbone)."""
    rng = np.random.default_rng(seed)
    u = rng.random(n_fam)
    p = np.where(u < 0.4, 0.97, np.where(u < 0.55, rng.uniform(0.2, 0.8, n_fam), 0.03))
    backbone = rng.permutation(n_fam)
    layouts = []
    for g in range(n_org):
        present = rng.random(n_fam) < p
        order = backbone.copy()
        for _ in range(int(rng.integers(0, 3))):
            a, b = sorted(rng.integers(0, n_fam, 2))
            order[a:b] = order[a:b][::-1].copy()
        lay = order[present[order]]
        layouts.append(lay.astype(np.int32))
    return layouts


def build_reference_pangenome(layouts, n_fam):
    from ppanggolin.pangenome import Pangenome
    from ppanggolin.genome import Organism, Contig, Gene
    from ppanggolin.geneFamily import GeneFamily
    pan = Pangenome()
    fams = [GeneFamily(i, f"fam{i}") for i in range(n_fam)]
    used = set(int(f) for lay in layouts for f in lay)
    for i in sorted(used):
        pan.add_gene_family(fams[i])
    for g, lay in enumerate(layouts):
        org = Organism(f"org{g}")
        ctg = Contig(g, f"ctg{g}", is_circular=True)
        org.add(ctg)
        pan.add_organism(org)
        genes = []
        for pos, f in enumerate(lay):
            gene = Gene(f"g{g}_{pos}")
            gene.fill_annotations(start=pos * 100 + 1, stop=pos * 100 + 90, strand="+", position=pos)
            gene.fill_parents(org, ctg)
            ctg.add(gene)
            fams[int(f)].add(gene)
            genes.append(gene)
        prev = None
        for gene in genes:
            if prev is not None:
                pan.add_edge(gene, prev)
            prev = gene
        if prev is not None and len(genes) > 0:
            pan.add_edge(genes[0], prev)
    for k in ("genomesAnnotated", "genesClustered", "neighborsGraph"):
        pan.status[k] = "Computed"
    return pan


def save_export_golden(ref_partition, pan):
    """What the reference's write_nem_input_files (partition.py:344-436) writes for the whole pangenome, parsed back."""
    with tempfile.TemporaryDirectory() as td:
        d = Path(td) / "exp"
        ref_partition.pan = pan
        ew, nb_fam = ref_partition.write_nem_input_files(d, set(pan.organisms), sm_degree=10)
        X = np.loadtxt(d / "nem_file.dat", dtype=np.uint8, ndmin=2)
        orgs = (d / "column_org_file").read_text().split()
        fams = [l.split("\t")[1].strip() for l in (d / "nem_file.index").read_text().splitlines()]
        nei = (d / "nem_file.nei").read_text().splitlines()[1:]
    np.savez_compressed(OUT / "export_fixedK.npz", X=X, orgs=np.array([o.strip('"') for o in orgs]), fams=np.array(fams),
                        nei=np.array(nei), edges_weight=ew, nb_fam=nb_fam)
    print("export golden", X.shape, ew, nb_fam)


def gbff_layouts(gbff_dir="/root/reference/testingDataset/GBFF"):
    """BASELINE config 1 surrogate: the Chlamydia trachomatis test genomes of the reference
    (GenBank flat files), CDS order per replicon, families = clusters of identical protein translations.  The real
    pipeline clusters with mmseqs2 (absent here), so the golden partition counts of expected_info_files/*.yaml do not
    apply; parity is ours vs the reference on the SAME surrogate pangenome.  Returns per genome a list of contigs, each a
    list of family ids in gene order."""
    import gzip
    import re
    from pathlib import Path as _P
    fam_of = {}
    genomes = []
    for f in sorted(_P(gbff_dir).glob("*.gbff.gz")):
        contigs, cur, in_tr, tr = [], None, False, []
        with gzip.open(f, "rt") as fh:
            for line in fh:
                if line.startswith("LOCUS"):
                    cur = []
                    contigs.append(cur)
                elif in_tr:
                    tr.append(line.strip())
                    if line.rstrip().endswith('"'):
                        in_tr = False
                        seq = "".join(tr).strip('"').replace(" ", "")
                        cur.append(fam_of.setdefault(seq, len(fam_of)))
                elif "/translation=" in line:
                    tr = [line.split("/translation=", 1)[1].strip()]
                    if tr[0].count('"') == 2:
                        seq = tr[0].strip('"')
                        cur.append(fam_of.setdefault(seq, len(fam_of)))
                    else:
                        in_tr = True
        genomes.append([c for c in contigs if len(c) > 0])
    return genomes, len(fam_of)


def build_reference_pangenome_contigs(genomes, n_fam):
    """Like build_reference_pangenome, several circular contigs per genome (graph/makeGraph.py:114-136 semantics)."""
    from ppanggolin.pangenome import Pangenome
    from ppanggolin.genome import Organism, Contig, Gene
    from ppanggolin.geneFamily import GeneFamily
    pan = Pangenome()
    fams = [GeneFamily(i, f"fam{i}") for i in range(n_fam)]
    used = sorted(set(int(f) for g in genomes for c in g for f in c))
    for i in used:
        pan.add_gene_family(fams[i])
    cid = 0
    for g, contigs in enumerate(genomes):
        org = Organism(f"org{g}")
        pan.add_organism(org)
        for ci, lay in enumerate(contigs):
            ctg = Contig(cid, f"ctg{g}_{ci}", is_circular=True)
            cid += 1
            org.add(ctg)
            genes = []
            for pos, f in enumerate(lay):
                gene = Gene(f"g{g}_{ci}_{pos}")
                gene.fill_annotations(start=pos * 100 + 1, stop=pos * 100 + 90, strand="+", position=pos)
                gene.fill_parents(org, ctg)
                ctg.add(gene)
                fams[int(f)].add(gene)
                genes.append(gene)
            prev = None
            for gene in genes:
                if prev is not None:
                    pan.add_edge(gene, prev)
                prev = gene
            if prev is not None and len(genes) > 0:
                pan.add_edge(genes[0], prev)
    for k in ("genomesAnnotated", "genesClustered", "neighborsGraph"):
        pan.status[k] = "Computed"
    return pan


def gbff_case(ref_partition):
    """partition() of the reference on the GBFF-derived surrogate pangenome: defaults of `ppanggolin all` (auto K over
    krange [3, 20], beta 2.5, max degree smoothing 10, chunk size 500, seed 42)."""
    genomes, n_fam = gbff_layouts()
    pan = build_reference_pangenome_contigs(genomes, n_fam)
    ref_partition.samples = []
    random.seed(1234)
    kw = dict(kval=-1, chunk_size=500, beta=2.5, free_dispersion=False, krange=[3, 20])
    with tempfile.TemporaryDirectory() as td:
        ref_partition.partition(pan, output=Path(td), tmpdir=Path(td), cpu=8, disable_bar=True, seed=42, **kw)
    labels = {fam.name: fam.partition for fam in pan.gene_families}
    names = sorted(labels, key=lambda s: int(s[3:]))
    params = pan.parameters["partition"]
    flat = np.concatenate([np.asarray(c, np.int32) for g in genomes for c in g])
    contig_ptr = np.cumsum([0] + [len(c) for g in genomes for c in g])
    genome_ptr = np.cumsum([0] + [len(g) for g in genomes])          # in contigs
    np.savez_compressed(OUT / "partition_gbff.npz", n_fam=n_fam, families=flat, contig_ptr=contig_ptr, genome_ptr=genome_ptr,
                        fam_names=np.array(names), labels=np.array([labels[n] for n in names]), kwargs=json.dumps(kw),
                        params=json.dumps(params))
    from collections import Counter
    print("gbff", len(genomes), "genomes", n_fam, "families", pan.number_of_edges, "edges",
          Counter(v[0] if v else "?" for v in labels.values()), params)


def partition_cases():
    ref_partition = import_reference_partition()
    gbff_case(ref_partition)
    cases = [("fixedK", dict(n_fam=900, n_org=60, seed=21), dict(kval=3, chunk_size=500, beta=2.5, free_dispersion=False)),
             ("fixedK_fd", dict(n_fam=900, n_org=60, seed=21), dict(kval=3, chunk_size=500, beta=2.6, free_dispersion=True)),
             ("autoK", dict(n_fam=700, n_org=50, seed=22), dict(kval=-1, chunk_size=500, beta=2.5, free_dispersion=False, krange=[3, 12])),
             ("chunked", dict(n_fam=500, n_org=90, seed=23), dict(kval=3, chunk_size=30, beta=2.5, free_dispersion=False))]
    for name, g, kw in cases:
        layouts = make_layouts(**g)
        pan = build_reference_pangenome(layouts, g["n_fam"])
        ref_partition.samples = []
        random.seed(1234)  # evaluate_nb_partitions' random.sample runs before partition() seeds
        with tempfile.TemporaryDirectory() as td:
            ref_partition.partition(pan, output=Path(td), tmpdir=Path(td), cpu=1, disable_bar=True, seed=42, **kw)
        if name == "fixedK":
            save_export_golden(ref_partition, pan)
        labels = {fam.name: fam.partition for fam in pan.gene_families}
        names = sorted(labels, key=lambda s: int(s[3:]))
        params = {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item()) for k, v in pan.parameters["partition"].items()}
        np.savez_compressed(OUT / f"partition_{name}.npz", n_fam=g["n_fam"], layouts=np.concatenate(layouts),
                            layout_ptr=np.cumsum([0] + [len(l) for l in layouts]), fam_names=np.array(names),
                            labels=np.array([labels[n] for n in names]), kwargs=json.dumps(kw), params=json.dumps(params))
        from collections import Counter
        print(name, Counter(v[0] if v else "?" for v in labels.values()), params)


def chunked_replay_cases():
    """Chunked partition() of the reference with every random.shuffle it performs recorded (genome names in shuffled order),
    so that a test can feed the SAME chunk samples, in the same rounds, to the GPU path (SURVEY.md section 7,
    "non-deterministic reference inputs"; partition.py:809-821)."""
    ref_partition = import_reference_partition()
    cases = [("chunked_replay", dict(n_fam=3000, n_org=150, seed=25), dict(kval=3, chunk_size=40, beta=2.5, free_dispersion=False)),
             ("chunked_replay_fd", dict(n_fam=2000, n_org=120, seed=26), dict(kval=3, chunk_size=30, beta=2.5, free_dispersion=True))]
    for name, g, kw in cases:
        layouts = make_layouts(**g)
        pan = build_reference_pangenome(layouts, g["n_fam"])
        ref_partition.samples = []
        recorded = []
        real_shuffle = random.shuffle

        def recording_shuffle(lst, *a):
            real_shuffle(lst, *a)
            recorded.append([o.name for o in lst])
        random.seed(1234)
        random.shuffle = recording_shuffle
        try:
            with tempfile.TemporaryDirectory() as td:
                ref_partition.partition(pan, output=Path(td), tmpdir=Path(td), cpu=1, disable_bar=True, seed=42, **kw)
        finally:
            random.shuffle = real_shuffle
        labels = {fam.name: fam.partition for fam in pan.gene_families}
        names = sorted(labels, key=lambda s: int(s[3:]))
        params = {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item()) for k, v in pan.parameters["partition"].items()}
        np.savez_compressed(OUT / f"partition_{name}.npz", n_fam=g["n_fam"], layouts=np.concatenate(layouts),
                            layout_ptr=np.cumsum([0] + [len(l) for l in layouts]), fam_names=np.array(names),
                            labels=np.array([labels[n] for n in names]), kwargs=json.dumps(kw), params=json.dumps(params),
                            shuffles=json.dumps(recorded), n_samples=len(ref_partition.samples))
        from collections import Counter
        print(name, Counter(v[0] if v else "?" for v in labels.values()), len(recorded), "shuffles", len(ref_partition.samples), "samples")


def rarefaction_case():
    """Reference raref_nem (rarefaction.py:36-189) on fixed genome samples of growing size: partition counts per sample."""
    ref_partition = import_reference_partition()
    import ppanggolin.nem.rarefaction as ref_raref
    g = dict(n_fam=700, n_org=50, seed=24)
    layouts = make_layouts(**g)
    pan = build_reference_pangenome(layouts, g["n_fam"])
    ref_partition.pan = pan
    orgs = {o.name: o for o in pan.organisms}
    rng = random.Random(7)
    names = sorted(orgs, key=lambda s: int(s[3:]))
    sample_names = [sorted(rng.sample(names, n)) for n in (12, 12, 20, 20, 35, 35, 50)]
    ref_raref.samples = [set(orgs[n] for n in sn) for sn in sample_names]
    out = []
    with tempfile.TemporaryDirectory() as td:
        for i in range(len(sample_names)):
            counts, idx = ref_raref.raref_nem(i, Path(td), beta=2.5, sm_degree=10, free_dispersion=False, chunk_size=500, kval=3,
                                              krange=[3, 20], seed=42)
            out.append(counts)
            print("raref", len(sample_names[i]), counts)
    np.savez_compressed(OUT / "rarefaction.npz", n_fam=g["n_fam"], layouts=np.concatenate(layouts),
                        layout_ptr=np.cumsum([0] + [len(l) for l in layouts]), samples=json.dumps(sample_names),
                        counts=json.dumps(out))


if __name__ == "__main__":
    refnem.build()
    if sys.argv[1:] == ["replay"]:
        chunked_replay_cases()
        sys.exit(0)
    nem_cases()
    partition_cases()
    chunked_replay_cases()
    rarefaction_case()
