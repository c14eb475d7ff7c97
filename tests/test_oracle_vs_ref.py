"""Live check of the oracle against oracle/_ref/libnem_ref.so (the unmodified reference C), where it has been built."""
import numpy as np
import pytest

from goldens import g6
from oracle import refnem
from ppanggolin_b200.synthetic import synth_pangenome

pytestmark = pytest.mark.skipif(not refnem.have_ref(), reason="oracle/_ref/libnem_ref.so not built (needs /root/reference)")


@pytest.mark.parametrize("N,D,K,beta,fd,itermax", [(700, 90, 3, 2.5, False, 100), (700, 90, 4, 2.5, True, 100),
                                                   (900, 130, 6, 0.0, False, 10), (500, 333, 2, 0.0, True, 10)])
def test_strict_oracle_equals_reference(tmp_path, N, D, K, beta, fd, itermax):
    sp = synth_pangenome(N, D, seed=N + D + K)
    be = sp.beta_eff(beta) if beta else 0.0
    rc, ref = refnem.ref_nem_run(sp.X, sp.rowptr, sp.col, sp.w, K, be, fd, itermax, workdir=tmp_path)
    o = refnem.oracle_nem_run(sp.X, sp.rowptr, sp.col, sp.w, K, be, fd, itermax)
    assert rc == 0 and o["status"] == 0
    _, T3, _ = refnem.oracle_labels(o["T"])
    assert np.array_equal(np.rint(T3 * 1000).astype(int), np.rint(ref["T3"] * 1000).astype(int))
    assert [g6(c) for c in o["crit"]] == [g6(c) for c in ref["crit"]]
    assert o["n_iter"] == ref["n_iter"]
