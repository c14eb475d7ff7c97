"""The oracle (oracle/nem_oracle.c, STRICT mode) against golden vectors produced by the unmodified reference C."""
import numpy as np
import pytest

from goldens import NEM_CASES, g6, load_nem
from oracle import refnem
from ppanggolin_b200.synthetic import unpack_rows


@pytest.mark.parametrize("name", NEM_CASES)
def test_oracle_strict_matches_reference_golden(name):
    g = load_nem(name)
    X = unpack_rows(g["bits"], g["D"])
    o = refnem.oracle_nem_run(X, g["rowptr"], g["col"], g["w"], g["K"], g["beta"], bool(g["fd"]), g["itermax"], mode=refnem.STRICT)
    if not g["ok"]:
        assert g["rc"] == 1 and o["status"] == 1   # EXIT_W_RESULT <-> empty class
        return
    assert o["status"] == 0
    assert o["n_iter"] == g["n_iter"] and o["converged"] == g["converged"]
    _, T3, _ = refnem.oracle_labels(o["T"])
    assert np.array_equal(np.rint(T3 * 1000).astype(int), np.rint(g["T3"] * 1000).astype(int))   # every " %5.3f " posterior
    assert [g6(c) for c in o["crit"]] == [g6(c) for c in g["crit"]]                                  # U D L M Z as "%g"
    assert np.array_equal(o["mu"], g["mu"])
    assert np.allclose([float("%5.3g" % v) for v in o["pk"]], g["pk"], rtol=0, atol=0)
    assert np.array_equal(np.array([[float("%10g" % v) for v in row] for row in o["eps"]]), g["eps"])


def test_appendix_b_known_answer():
    """the published known-answer vector of the compiled reference."""
    g = load_nem("appendixB")
    assert [g6(c) for c in g["crit"]] == [-3599.46, -3749.47, -3749.47, -4129.86, -680.411]
    assert g["n_iter"] == 2
    rows, counts = np.unique(np.rint(g["T3"] * 1000).astype(int), axis=0, return_counts=True)
    hist = {tuple(r): c for r, c in zip(rows, counts)}
    assert hist[(1000, 0, 0)] == 127 and hist[(0, 0, 1000)] == 106 and hist[(0, 1000, 0)] == 66 and hist[(0, 11, 989)] == 1


@pytest.mark.parametrize("name", ["sk_graph", "skd_graph", "sweep_K5", "sweep_fd_K7", "wide1000"])
def test_oracle_fp64_within_parity_bar(name):
    """The closed-form fp64 restatement (what the CUDA kernels compute) stays inside the north-star bar against the
    reference: labels >= 99.9 %, criterion M within 1e-5 relative."""
    g = load_nem(name)
    X = unpack_rows(g["bits"], g["D"])
    o = refnem.oracle_nem_run(X, g["rowptr"], g["col"], g["w"], g["K"], g["beta"], bool(g["fd"]), g["itermax"], mode=refnem.FP64)
    lab, _, _ = refnem.oracle_labels(o["T"])
    ref_lab, _, _ = refnem.oracle_labels(g["T3"].astype(np.float32))
    assert (lab == ref_lab).mean() >= 0.999
    assert abs(o["crit"][3] - g["crit"][3]) <= 1e-5 * abs(g["crit"][3])


def test_gauss_seidel_is_detectable():
    """Appendix B: a parallel (Jacobi) E-step takes 3 iterations and lands on a different M."""
    g = load_nem("appendixB")
    X = unpack_rows(g["bits"], g["D"])
    o = refnem.oracle_nem_run(X, g["rowptr"], g["col"], g["w"], 3, g["beta"], False, 100, jacobi=True)
    assert o["n_iter"] == 3 and g6(o["crit"][3]) != g6(g["crit"][3])
