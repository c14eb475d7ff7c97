"""Parity at the shapes BASELINE.json names (configs[1], [2], [3] and a wide un-chunked matrix), against golden vectors made
with the UNMODIFIED reference C where it is numerically valid (D <= 1000) and with the fp64 restatement for the wide
matrix (tests/golden/make_golden_configs.py).  The instances are regenerated from their seeds and checked against the
digest stored with the golden before anything is compared.

Bars (BASELINE.json north_star): labels identical for >= 99.9 % of families, disagreements only at posterior near-ties;
criterion M and ICL within 1e-5 relative; the same chosen K.  Iteration counts are reported, not asserted.
"""
import numpy as np
import pytest

from goldens import REPLAY_CASES, g6, instance_digest, load_config, load_partition
from ppanggolin_b200.synthetic import synth_pangenome, synth_pangenome_blocked

pytestmark = pytest.mark.gpu

REL = 1e-5


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from ppanggolin_b200.nem import engine
    engine.lib()
    return engine


def _near_tie(row):
    s = np.sort(row)[::-1]
    return (s[0] - s[1]) <= 0.002 or abs(s[0] - 0.5) <= 0.002


def _check_labels(lab, ref_lab, T):
    diff = np.flatnonzero(lab != ref_lab)
    assert len(diff) <= 0.001 * len(lab), f"{len(diff)} of {len(lab)} labels differ"
    for i in diff:
        assert _near_tie(T[i]), f"row {i} differs away from a tie: {T[i]}"
    return len(diff)


def _mu_bits(mu):
    return np.packbits(mu == 1.0, axis=1)


def test_cfg2_single_run_matches_reference(eng):
    """configs[1]: 20 000 families x 1 000 genomes, K = 3, beta 2.5, sk_ -- the reference's own run of this instance."""
    g = load_config("cfg2_ref")
    sp = synth_pangenome(g["N"], g["D"], seed=42)
    assert instance_digest(sp) == g["digest"]
    inp = eng.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    run = eng.nem_run_device(inp, 3, float(g["beta"]))
    assert run.status == 0 and g["rc"] == 0
    lab, _ = eng.labels_device(run.T)
    T = run.T.cpu().numpy()
    ndiff = _check_labels(lab.cpu().numpy(), g["labels"].astype(np.int32), T)
    assert abs(run.crit[3] - g["crit"][3]) <= REL * abs(g["crit"][3])
    assert abs(run.crit[1] - g["crit"][1]) <= REL * abs(g["crit"][1])
    T3 = np.rint(T.astype(np.float64) * 1000).astype(np.int64)
    moved = (np.abs(T3 - g["T3"].astype(np.int64)).max(1) > 0).mean()
    assert moved <= 0.001                                     # third printed decimal of the posteriors
    assert np.array_equal(_mu_bits(run.mu), g["mu"]) and int((run.mu == 0.5).sum()) == g["mu_half"]
    assert np.allclose(run.pk, g["pk"], atol=6e-4)            # the golden holds the " %5.3g " text of the .mf file
    print(f"cfg2: iterations ours/reference {run.n_iter}/{g['n_iter']}, labels differing {ndiff}, posterior rows moved {moved:.5f}, "
          f"rel M {abs(run.crit[3] - g['crit'][3]) / abs(g['crit'][3]):.2e}")


def test_cfg3_k_sweep_matches_reference(eng):
    """configs[2], K selection: 50 000 families x 500 sampled genomes, beta 0, 10 iterations, K = 2..20 as ONE batch; the
    reference ran every K on the same matrix.  M and ICL per K within 1e-5, and the SAME chosen K."""
    from ppanggolin_b200.nem import partition as part
    g = load_config("cfg3_ref")
    sp = synth_pangenome(g["N"], g["D"], seed=43)
    assert instance_digest(sp) == g["digest"]
    inp = eng.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    ks = [int(k) for k in g["ks"]]
    runs = eng.nem_run_batch([(inp, k, 0.0, False, 10) for k in ks])
    ours, ref, iters = [], [], []
    for k, run in zip(ks, runs):
        r = g["summary"][k]
        assert run.status == 0 and r["ok"] == 1
        lab, ent = eng.labels_device(run.T)
        assert abs(run.crit[3] - r["M"]) <= REL * abs(r["M"]), (k, run.crit[3], r["M"])
        ours.append((k, g6(run.crit[3]), ent))
        ref.append((k, r["M"], r["entropy"]))
        iters.append((k, run.n_iter, r["n_iter"]))
        if f"labels_K{k}" in g:   # not an output of these runs (only M and the entropy are), kept as a sanity bar
            assert (lab.cpu().numpy() == g[f"labels_K{k}"].astype(np.int32)).mean() >= 0.995
    k_ours, _, icl_o, _ = part.select_k(ours, g["D"], g["N"], False, 0.05)
    k_ref, _, icl_r, _ = part.select_k(ref, g["D"], g["N"], False, 0.05)
    for k in ks:
        assert abs(icl_o[k] - icl_r[k]) <= REL * abs(icl_r[k]), (k, icl_o[k], icl_r[k])
    assert k_ours == k_ref
    print("cfg3: chosen K", k_ours, "iterations (K, ours, reference) that differ:", [t for t in iters if t[1] != t[2]])


def test_cfg4_chunk_instance_matches_reference(eng):
    """configs[3], one chunk instance: 100 000 families x 500 genomes, K = 3, beta 2.5, free dispersion (skd)."""
    g = load_config("cfg4_ref")
    sp = synth_pangenome(g["N"], g["D"], seed=44)
    assert instance_digest(sp) == g["digest"]
    inp = eng.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    run = eng.nem_run_device(inp, 3, float(g["beta"]), True)
    assert run.status == 0 and g["rc"] == 0
    lab, _ = eng.labels_device(run.T)
    T = run.T.cpu().numpy()
    ndiff = _check_labels(lab.cpu().numpy(), g["labels"].astype(np.int32), T)
    assert abs(run.crit[3] - g["crit"][3]) <= REL * abs(g["crit"][3])
    T3 = np.rint(T.astype(np.float64) * 1000).astype(np.int64)
    moved = (np.abs(T3 - g["T3"].astype(np.int64)).max(1) > 0).mean()
    assert moved <= 0.001
    assert (_mu_bits(run.mu) != g["mu"]).sum() == 0
    print(f"cfg4: iterations ours/reference {run.n_iter}/{g['n_iter']}, labels differing {ndiff}, posterior rows moved {moved:.5f}, "
          f"rel M {abs(run.crit[3] - g['crit'][3]) / abs(g['crit'][3]):.2e}")


def test_wide_matrix_matches_fp64_oracle(eng):
    """20 000 families x 50 000 genomes, un-chunked (the regime of configs[4]): the reference underflows there (SURVEY.md
    section 0-5), so the golden comes from the fp64 restatement; log-space epilogue on the GPU."""
    g = load_config("wide_fp64")
    sp = synth_pangenome_blocked(g["N"], g["D"], seed=45)
    assert instance_digest(sp) == g["digest"]
    inp = eng.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    run = eng.nem_run_device(inp, 3, float(g["beta"]), flags=eng.EPI_LOGSPACE)
    assert run.status == 0 and g["status"] == 0
    lab, ent = eng.labels_device(run.T)
    T = run.T.cpu().numpy()
    ndiff = _check_labels(lab.cpu().numpy(), g["labels"].astype(np.int32), T)
    assert abs(run.crit[3] - g["crit"][3]) <= REL * abs(g["crit"][3])
    assert abs(run.crit[2] - g["crit"][2]) <= REL * abs(g["crit"][2])
    T3 = np.rint(T.astype(np.float64) * 1000).astype(np.int64)
    moved = (np.abs(T3 - g["T3"].astype(np.int64)).max(1) > 0).mean()
    assert moved <= 0.001
    assert (_mu_bits(run.mu) != g["mu"]).mean() < 1e-4
    print(f"wide: iterations ours/oracle {run.n_iter}/{g['n_iter']}, labels differing {ndiff}, rows moved {moved:.5f}")


@pytest.mark.parametrize("name", REPLAY_CASES)
def test_chunked_partition_with_the_reference_samples(eng, name, tmp_path, monkeypatch):
    """Chunked partition() (partition.py:781-864) fed with the SAME chunk samples, in the same rounds, as the reference run
    that produced the golden (every random.shuffle of that run was recorded): >= 99.9 % of the final labels identical."""
    import random
    import minipan
    from ppanggolin_b200.nem import partition as part
    g = load_partition(name)
    pan = minipan.build_from_layouts(g["layouts"], g["n_fam"])
    by_name = {o.name: o for o in pan.organisms}
    replay = iter(g["shuffles"])
    extra = []

    def replay_shuffle(lst, *a):
        try:
            order = next(replay)
        except StopIteration:           # the reference needed fewer rounds: fall back to a real shuffle and report it
            extra.append(1)
            random.Random(len(extra)).shuffle(lst)
            return
        assert sorted(o.name for o in lst) == sorted(order)
        lst[:] = [by_name[n] for n in order]
    monkeypatch.setattr(random, "shuffle", replay_shuffle)
    part.partition(pan, output=tmp_path, tmpdir=tmp_path, cpu=1, disable_bar=True, seed=42, **g["kwargs"])
    ours = {f.name: f.partition for f in pan.gene_families}
    ref = dict(zip(g["fam_names"], g["labels"]))
    assert sorted(ours) == sorted(ref)
    agree = np.mean([ours[n] == ref[n] for n in ref])
    assert not extra and len(part.samples) == g["n_samples"], (len(extra), len(part.samples), g["n_samples"])
    assert agree >= 0.999, agree
    for k, v in g["params"].items():
        assert pan.parameters["partition"][k] == v, k
