"""export_arrays.PangenomeArrays.instance == export.export_from_pangenome (the object traversal that mirrors
partition.py:344-436, itself pinned to the reference's text export in test_host_logic.py) for the whole pangenome and
for random genome subsets, including self-loop edges, the sm_degree cut and the zero-weight behaviour of the C reader."""
import random

import numpy as np
import pytest

import minipan
from goldens import load_partition
from ppanggolin_b200.nem import export
from ppanggolin_b200.nem.export_arrays import PangenomeArrays
from ppanggolin_b200.synthetic import unpack_rows


def _same_instance(a, b):
    assert a.fam_names == b.fam_names and a.D == b.D
    assert a.edges_weight == pytest.approx(b.edges_weight, rel=1e-12)
    Xa = unpack_rows(a.host_bits(), a.D)[:, np.argsort(a.org_names)]
    Xb = unpack_rows(b.host_bits(), b.D)[:, np.argsort(b.org_names)]
    assert np.array_equal(Xa, Xb)
    assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col) and np.array_equal(a.w, b.w)
    assert int(np.unpackbits(a.host_bits().view(np.uint8), axis=1, bitorder="little")[:, a.D:].sum()) == 0


@pytest.mark.parametrize("case", ["fixedK", "chunked"])
def test_instances_match_object_traversal(case):
    g = load_partition(case)
    pan = minipan.build_from_layouts(g["layouts"], g["n_fam"])
    arrays = PangenomeArrays.from_pangenome(pan, "cpu")
    orgs = list(pan.organisms)
    rnd = random.Random(3)
    subsets = [set(orgs)] + [set(rnd.sample(orgs, n)) for n in (5, 17, 30, len(orgs) - 1)]
    for sub in subsets:
        for sm in (10, 4):
            _same_instance(arrays.instance(sub, sm), export.export_from_pangenome(pan, sub, sm))


def test_duplicates_self_loops_and_tiny_weights():
    """Tandem duplicates give self-loop edges and multi-pair counts (makeGraph.py:122-126); with > 20 000 genomes a single
    pair rounds to weight 0.0 and the C reader's index/weight misalignment appears (nem_exe.c:1413-1451)."""
    rng = np.random.default_rng(0)
    n_fam, n_org = 60, 12
    layouts = []
    for g_ in range(n_org):
        lay = [int(f) for f in rng.permutation(n_fam)[: rng.integers(20, 50)]]
        for _ in range(5):      # tandem duplicates
            p = int(rng.integers(0, len(lay)))
            lay.insert(p, lay[p])
        layouts.append(np.array(lay))
    pan = minipan.build_from_layouts(layouts, n_fam)
    arrays = PangenomeArrays.from_pangenome(pan, "cpu")
    orgs = list(pan.organisms)
    for sub in (set(orgs), set(orgs[:5]), set(orgs[3:9])):
        _same_instance(arrays.instance(sub, 10), export.export_from_pangenome(pan, sub, 10))
    # zero-weight quirk: pretend the sample has 30 000 genomes by padding empty organisms
    for i in range(30_000 - n_org):
        o = minipan.Organism(f"pad{i}")
        pan._orgs[o.name] = o
    arrays = PangenomeArrays.from_pangenome(pan, "cpu")
    allorgs = set(pan.organisms)
    a, b = arrays.instance(allorgs, 10), export.export_from_pangenome(pan, allorgs, 10)
    assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col) and np.array_equal(a.w, b.w)
    assert a.edges_weight == pytest.approx(b.edges_weight, rel=1e-12)
    assert len(b.w) < sum(len(list(f.edges)) for f in pan.gene_families)   # some weights were dropped
