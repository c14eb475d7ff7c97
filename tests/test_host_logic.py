"""Host-side logic of the drop-in module: in-memory export vs the reference's text export, K selection, sharding."""
import json
import math

import numpy as np
import pytest

import minipan
from goldens import GOLD, load_partition
from ppanggolin_b200.nem import export, partition as part
from ppanggolin_b200.nem.sharded import shard_rows
from ppanggolin_b200.synthetic import pack_rows, unpack_rows


def test_export_matches_reference_text_export():
    """export_from_pangenome == what partition.py:344-436 writes (golden parsed from the reference's files)."""
    g = load_partition("fixedK")
    z = np.load(GOLD / "export_fixedK.npz")
    pan = minipan.build_from_layouts(g["layouts"], g["n_fam"])
    inp = export.export_from_pangenome(pan, set(pan.organisms), sm_degree=10, keep_text=True)
    assert inp.N == int(z["nb_fam"]) and inp.D == z["X"].shape[1]
    assert inp.edges_weight == pytest.approx(float(z["edges_weight"]), rel=1e-12)
    # rows in the same family order; columns compared after sorting by genome name (set order is run-dependent)
    assert inp.fam_names == [str(s) for s in z["fams"]]
    X = unpack_rows(inp.bits, inp.D)
    ours = X[:, np.argsort(inp.org_names)]
    ref = z["X"][:, np.argsort([str(s) for s in z["orgs"]])]
    assert np.array_equal(ours, ref)
    # neighbour lists: same text rows as the .nei file
    for i, line in enumerate(z["nei"]):
        f = str(line).split("\t")
        assert int(f[0]) == i + 1
        n = int(f[1])
        ent = inp.nei_text[i]
        if n == 0:
            assert ent is None
        else:
            idx, wt = ent
            assert [int(v) - 1 for v in f[2:2 + n]] == idx and f[2 + n:] == wt
            a, b = inp.rowptr[i], inp.rowptr[i + 1]
            assert list(inp.col[a:b]) == idx and np.array_equal(inp.w[a:b], np.array([float(t) for t in wt], np.float32))
    assert int(np.unpackbits(inp.bits.view(np.uint8), axis=1, bitorder="little")[:, inp.D:].sum()) == 0   # padding bits are zero


def test_c_reader_zero_weight_quirk():
    """nem_exe.c:1413-1451: zero weights are skipped after the indices were stored -> first nv indices, non-zero weights."""
    idx, kept = export.c_reader_neighbours([5, 7, 9], ["0.1", "0.0", "0.3"])
    assert idx == [5, 7] and [float(v) for v in kept] == [np.float32(0.1), np.float32(0.3)]


def test_select_k_follows_reference_rule():
    # (K, M, entropy); BIC = M - 0.5 ln(nb_params) nb_fam; ICL = BIC - entropy (partition.py:516-545)
    runs = [(2, -1000.0, -1.0), (3, -900.0, -2.0), (4, -890.0, -3.0), (5, -889.0, -30.0), (6, None, None)]
    k, bics, icls, lls = part.select_k(runs, nb_org=10, nb_fam=50, free_dispersion=False, icl_margin=0.05)
    for kk, m, e in runs[:4]:
        assert bics[kk] == pytest.approx(m - 0.5 * math.log(kk * 12) * 50)
        assert icls[kk] == pytest.approx(bics[kk] - e)
    best = max(icls, key=icls.get)
    delta = (icls[best] - min(icls.values())) * 0.05
    want = min(kk for kk, v in icls.items() if v >= icls[best] - delta and kk <= best)
    assert k == (want if want >= 3 else 3)
    assert part.select_k(runs[:3], 10, 50, False, 0.05)[0] == 3      # needs > 3 successful candidates


def test_shard_rows_partition():
    for N in (1, 7, 200_000, 200_003):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(N, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    for D in (1, 31, 32, 33, 127, 128, 129, 500, 1000):
        X = (rng.random((17, D)) < 0.4).astype(np.uint8)
        b = pack_rows(X)
        assert b.shape[1] % 4 == 0 and b.shape[1] * 32 >= D
        assert np.array_equal(unpack_rows(b, D), X)
        assert int(np.unpackbits(b.view(np.uint8), axis=1, bitorder="little")[:, D:].sum()) == 0


def test_row_pitch_rules():
    """The C ABI takes any multiple of 4 words that holds D bits; the exporters pad wide matrices to whole 128-byte lines
    (the streaming kernels' cp.async copies need rows that start on a sector boundary) and keep narrow ones compact."""
    from ppanggolin_b200.synthetic import row_stride_words
    for D in (1, 31, 32, 33, 127, 128, 129, 500, 1000, 2047, 2048, 2049, 4100, 20000, 50000, 99999):
        W = row_stride_words(D)
        assert W * 32 >= D and W % 4 == 0
        words = (D + 31) // 32
        if words <= 64:
            assert W < words + 4               # at most 3 padding words
        else:
            assert W % 32 == 0 and W < words + 32
    assert row_stride_words(500) == 16 and row_stride_words(50000) == 1568


def test_parser_flags_match_reference_defaults():
    import argparse
    p = argparse.ArgumentParser()
    part.parser_partition(p)
    a = p.parse_args([])
    assert (a.beta, a.max_degree_smoothing, a.free_dispersion, a.chunk_size, a.nb_of_partitions, a.krange, a.ICL_margin,
            a.seed, a.cpu, a.draw_ICL, a.keep_tmp_files) == (2.5, 10, False, 500, -1, [3, 20], 0.05, 42, 1, False, False)
    a = p.parse_args("-b 2.6 -ms 10 -fd -ck 500 -Kmm 3 12 -im 0.04 --draw_ICL".split())
    assert a.max_degree_smoothing == 10.0 and isinstance(a.max_degree_smoothing, float) and a.krange == [3, 12]


def test_run_partitioning_requires_export_and_rejects_random_init(tmp_path):
    with pytest.raises(FileNotFoundError):
        part.run_partitioning(tmp_path / "nothing", 10)
    with pytest.raises(NotImplementedError):
        part.run_partitioning(tmp_path, 10, init="random")


def test_rarefaction_parser_flags_match_reference_defaults():
    """parser_rarefaction (rarefaction.py:775-912): same flags, defaults and types as the reference."""
    import argparse
    from pathlib import Path
    from ppanggolin_b200.nem import rarefaction as raref
    ap = argparse.ArgumentParser()
    raref.parser_rarefaction(ap)
    a = ap.parse_args([])
    assert (a.beta, a.depth, a.min, a.max, a.max_degree_smoothing) == (2.5, 30, 1, 100, 10)
    assert isinstance(a.max, (int, float)) and isinstance(ap.parse_args(["--max", "7"]).max, float)
    assert (a.free_dispersion, a.chunk_size, a.nb_of_partitions, a.reestimate_K, a.krange) == (False, 500, -1, False, [3, -1])
    assert (a.soft_core, a.seed, a.cpu) == (0.95, 42, 1) and a.pangenome is None
    b = ap.parse_args(["-p", "x.h5", "-b", "3", "--depth", "5", "-ms", "4", "-fd", "-ck", "50", "-K", "4", "--reestimate_K",
                       "-Kmm", "3", "9", "--soft_core", "0.9", "-se", "7", "-c", "3"])
    assert b.pangenome == Path("x.h5") and b.krange == [3, 9] and b.free_dispersion and b.reestimate_K and b.max_degree_smoothing == 4.0
    sub = argparse.ArgumentParser().add_subparsers()
    assert raref.subparser(sub).prog.endswith("rarefaction")
