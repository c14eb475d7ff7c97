"""The CUDA exporter (nemb200_export_samples, csrc/nemb200_export.cuh) against the object traversal that follows the
reference's write_nem_input_files line by line (export.export_from_pangenome; itself pinned to the reference's text
export in tests/test_host_logic.py), plus the device-side chunk votes and subset counts."""
import random

import numpy as np
import pytest

import minipan
import npref
from goldens import load_partition
from ppanggolin_b200.nem import export

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from ppanggolin_b200.nem import engine
    engine.lib()
    return engine


def _pan(name="fixedK"):
    g = load_partition(name)
    return minipan.build_from_layouts(g["layouts"], g["n_fam"])


def _same(inst, dp, ref, eng, check_levels=True):
    """DeviceInstance == NemInput of the object traversal (rows in family order, columns in ascending genome order)."""
    assert inst.N == ref.N and inst.D == ref.D
    assert [dp.fam_names[i] for i in inst.fam_rows.cpu().numpy()] == ref.fam_names
    assert np.array_equal(inst.inp.bits.cpu().numpy().view(np.uint32), ref.bits)
    assert inst.edges_weight == pytest.approx(ref.edges_weight, rel=1e-12)
    nnz = int(ref.rowptr[-1])
    if nnz == 0:
        assert inst.inp.rowptr is None
        return
    assert np.array_equal(inst.inp.rowptr.cpu().numpy(), ref.rowptr)
    assert np.array_equal(inst.inp.col.cpu().numpy(), ref.col)
    assert np.array_equal(inst.inp.w.cpu().numpy().view(np.uint32), ref.w.view(np.uint32))     # bit-identical float32 weights
    if check_levels:
        lp, rows, nl = eng.gs_schedule(ref.rowptr, ref.col)
        assert inst.inp.n_levels == nl
        glp = inst.inp.sched_ptr.cpu().numpy()
        grows = inst.inp.sched_rows.cpu().numpy()
        assert np.array_equal(glp, lp)
        for l in range(nl):   # rows of one level are independent: any order inside a level is a valid schedule
            assert np.array_equal(np.sort(grows[glp[l]:glp[l + 1]]), np.sort(rows[lp[l]:lp[l + 1]]))


def _ordered_export(pan, dp, sub, sm):
    """export_from_pangenome with the genome columns in ascending column order (the order the exporter uses)."""
    orgs = sorted(sub, key=lambda o: dp.org_index[o])

    class _Ordered(set):      # a set that iterates in column order: column order = iteration order (partition.py:377-379)
        def __iter__(self):
            return iter(orgs)
    return export.export_from_pangenome(pan, _Ordered(sub), sm)


@pytest.mark.parametrize("sm", [10, 3, 4.5])
def test_batch_export_equals_object_traversal(eng, sm):
    from ppanggolin_b200.nem.export_device import DevicePangenome
    pan = _pan()
    dp = DevicePangenome.from_pangenome(pan)
    orgs = list(pan.organisms)
    rng = random.Random(5)
    for size in (1, 7, 25, len(orgs)):
        subs = [set(rng.sample(orgs, size)) for _ in range(5)]
        insts = dp.export([dp.columns(s) for s in subs], sm)
        for s, inst in zip(subs, insts):
            _same(inst, dp, _ordered_export(pan, dp, s, sm), eng)
    # all genomes, identity selection (no gather): the same instance
    allg = set(orgs)
    _same(dp.export([None], sm)[0], dp, _ordered_export(pan, dp, allg, sm), eng)


def test_export_of_multi_contig_genomes_and_zero_weight_quirk(eng):
    """Several contigs per genome (tandem self-loops, gene pairs > 1 per genome) and weights that print as 0.0 (dropped by the
    C reader after the indices were stored, nem_exe.c:1413-1451): 12 000 copies of a genome set make pairs / n < 5e-5."""
    from goldens import load_partition_gbff
    from ppanggolin_b200.nem.export_device import DevicePangenome
    g = load_partition_gbff()
    pan = minipan.build_from_contigs(g["genomes"], g["n_fam"])
    dp = DevicePangenome.from_pangenome(pan)
    orgs = list(pan.organisms)
    rng = random.Random(9)
    subs = [set(rng.sample(orgs, 20)) for _ in range(3)]
    for s, inst in zip(subs, dp.export([dp.columns(s) for s in subs], 10)):
        _same(inst, dp, _ordered_export(pan, dp, s, 10), eng)
    _same(dp.export([None], 10)[0], dp, _ordered_export(pan, dp, set(orgs), 10), eng)


def test_copresence_model_equals_explicit_table(eng):
    """The synthetic-at-scale edge model (one gene pair in genome g iff both families are present in g) against the same
    pangenome given as an explicit (edge, genome, pairs) table."""
    import torch
    from ppanggolin_b200.nem.export_device import DevicePangenome
    from ppanggolin_b200.synthetic import pack_rows, row_stride_words
    rng = np.random.default_rng(3)
    F, G = 900, 130
    X = (rng.random((F, G)) < rng.random(F)[:, None]).astype(np.uint8)
    X[X.sum(1) == 0, 0] = 1
    edges = np.unique(np.sort(rng.integers(0, F, (2500, 2)), axis=1), axis=0)
    edges = edges[edges[:, 0] != edges[:, 1]]
    bits = torch.from_numpy(pack_rows(X).view(np.int32).reshape(F, row_stride_words(G))).cuda()
    dp_a = DevicePangenome.synthetic_copresence(bits, G, edges)
    both = (X[edges[:, 0]] & X[edges[:, 1]]).astype(bool)                      # [E, G]
    ptr = np.concatenate([[0], np.cumsum(both.sum(1))]).astype(np.int32)
    org = np.nonzero(both)[1].astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.int32)).cuda()
    dp_b = DevicePangenome(bits, G, dp_a.adj_ptr, dp_a.adj_other, dp_a.adj_edge, len(edges), cov=(t(ptr), t(org), t(np.ones(len(org)))))
    subs = [sorted(rng.choice(G, 40, replace=False).tolist()) for _ in range(4)]
    for a, b in zip(dp_a.export(subs, 10), dp_b.export(subs, 10)):
        assert a.N == b.N and a.edges_weight == b.edges_weight and a.inp.n_levels == b.inp.n_levels
        for x, y in ((a.inp.bits, b.inp.bits), (a.fam_rows, b.fam_rows), (a.inp.rowptr, b.inp.rowptr), (a.inp.col, b.inp.col),
                     (a.inp.w, b.inp.w), (a.inp.sched_ptr, b.inp.sched_ptr)):
            assert torch.equal(x, y)


def test_chunk_votes_kernel_equals_reference_rule(eng):
    import torch
    from ppanggolin_b200.nem.export_device import ChunkVotes
    rng = np.random.default_rng(8)
    F, n_org, chunk = 700, 40, 10
    v = ChunkVotes(F, n_org, chunk, torch.device("cuda"))
    votes = np.zeros((F, 4), np.int64)
    validated = np.zeros(F, bool)
    for rnd in range(6):
        codes = np.where(rng.random((5, F)) < 0.7, rng.integers(0, 3, (5, F)), -1).astype(np.int8)
        npref.apply_votes_ref(codes, votes, validated, n_org, chunk)
        v.apply(torch.from_numpy(codes).cuda())
        assert np.array_equal(v.votes.cpu().numpy(), votes) and np.array_equal(v.validated.cpu().numpy().astype(bool), validated)
        assert v.count_validated() == int(validated.sum())
    # label rule of partition.py:206-231 -> vote letters, scattered through fam_rows
    K, N = 4, 300
    T = rng.random((N, K)).astype(np.float32); T /= T.sum(1, keepdims=True)
    T[:40] = 0.25                                             # ties -> "S_"
    T[40:80] = np.eye(K, dtype=np.float32)[rng.integers(0, K, 40)]
    fam_rows = rng.choice(F, N, replace=False).astype(np.int32)
    out = torch.full((F,), -1, dtype=torch.int8, device="cuda")
    v.codes(torch.from_numpy(T).cuda(), K, torch.from_numpy(fam_rows).cuda(), out)
    from oracle import refnem
    lab, _, _ = refnem.oracle_labels(T)
    want = np.full(F, -1, np.int8)
    want[fam_rows] = np.where(lab == 0, 0, np.where(lab == K - 1, 2, 1))
    assert np.array_equal(out.cpu().numpy(), want)


def test_subset_counts_equal_numpy(eng):
    from ppanggolin_b200.nem.export_device import DevicePangenome
    pan = _pan()
    dp = DevicePangenome.from_pangenome(pan)
    orgs = list(pan.organisms)
    rng = random.Random(2)
    fams = list(pan.gene_families)
    for size in (1, 13, len(orgs)):
        sub = set(rng.sample(orgs, size))
        got = dp.subset_counts(dp.columns(sub)).cpu().numpy()
        want = np.array([len(sub & set(f.organisms)) for f in fams])
        assert np.array_equal(got, want)
