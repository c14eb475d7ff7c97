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


def test_graph_built_on_arrays_equals_object_graph():
    """graph_arrays.build_from_contigs (makeGraph.py:87-138 on arrays: consecutive genes, circular closure, tandem self-loops,
    several contigs per genome) gives the arrays one traversal of the object graph gives: same instances for any subset."""
    from goldens import load_partition_gbff
    from ppanggolin_b200.nem import graph_arrays
    g = load_partition_gbff()
    genomes = [[np.asarray(c) for c in contigs] for contigs in g["genomes"][:20]]
    pan = minipan.build_from_contigs(genomes, g["n_fam"])
    ref = PangenomeArrays.from_pangenome(pan, "cpu")
    got = graph_arrays.build_from_contigs(genomes, g["n_fam"])
    assert got.fam_names == ref.fam_names and got.n_edges == ref.n_edges
    assert np.array_equal(got.presence.numpy(), ref.presence.numpy())
    assert np.array_equal(got.adj_ptr.numpy(), ref.adj_ptr.numpy()) and np.array_equal(got.adj_other.numpy(), ref.adj_other.numpy())
    orgs = list(pan.organisms)
    rnd = random.Random(4)
    for size in (len(orgs), 9, 3):
        sub = rnd.sample(range(len(orgs)), size)
        a = got.instance(sub, 10)
        b = ref.instance({orgs[i] for i in sub}, 10)
        assert a.fam_names == b.fam_names and a.edges_weight == pytest.approx(b.edges_weight, rel=1e-12)
        assert np.array_equal(a.host_bits(), b.host_bits())
        assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col) and np.array_equal(a.w, b.w)


def test_graph_on_arrays_fragments_and_copy_number_cut():
    """The two rules the object builders of the tests do not exercise: a pair of the same family is not linked when one of
    the genes is a fragment (makeGraph.py:116-121), and families with too many copies in a genome are left out (:66-75)."""
    from ppanggolin_b200.nem import graph_arrays
    genomes = [[np.array([0, 1, 1, 2, 3])], [np.array([0, 2, 2, 2, 3])]]
    frags = [[np.array([False, False, True, False, False])], [np.array([False] * 5)]]
    a = graph_arrays.build_from_contigs(genomes, 4, fragments=frags, circular=False)
    # genome 0: 0-1, (1,1 skipped: fragment), 1-2, 2-3; genome 1: 0-2, 2-2, 2-2, 2-3  -> edges {01, 12, 23, 02, 22}
    assert a.n_edges == 5
    pairs = dict(zip(zip(a.cov_edge.tolist(), a.cov_org.tolist()), a.cov_cnt.tolist()))
    e22 = [e for e in range(5) if (e, 1) in pairs and pairs[(e, 1)] == 2]
    assert len(e22) == 1                                            # the 2-2 self-loop holds two gene pairs in genome 1
    b = graph_arrays.build_from_contigs(genomes, 4, fragments=frags, circular=False, remove_copy_number=3)
    # family 2 has 3 copies in genome 1: removed everywhere -> genome 0: 0-1, 1-3 (prev skips over 2); genome 1: 0-3
    assert b.n_edges == 3
