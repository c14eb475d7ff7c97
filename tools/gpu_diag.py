"""First-contact GPU diagnostic: every kernel step against numpy, then whole runs against the oracle.
Run on the GPU box:  python tools/gpu_diag.py > gpurun_out/diag.log 2>&1"""
import sys, time, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import npref
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome
from oracle import refnem


def steps(N, D, K, fd, seed=3):
    print(f"--- steps N={N} D={D} K={K} fd={fd}")
    sp = synth_pangenome(N, D, seed)
    inp = engine.DeviceInput.from_numpy(sp.bits, D, sp.rowptr, sp.col, sp.w)
    Ws = sp.bits.shape[1]
    plan = engine.Plan(N, N, D, Ws, K, fd, engine.EPI_REF)
    plan.init_params()
    mu, eps, pk, st = plan.get_params()
    pk0, mu0, eps0 = npref.init_params(K, D)
    print(" init: mu", np.array_equal(mu, mu0), "eps", np.array_equal(eps, eps0), "pk", np.array_equal(pk, pk0), "status", st)
    logf = torch.empty((N, K), dtype=torch.float64, device="cuda")
    plan.density(inp.bits, logf)
    lf = logf.cpu().numpy(); lf0 = npref.logf(sp.X, mu0, eps0)
    print(" density(init): max abs err", np.abs(lf - lf0).max(), "max |lf|", np.abs(lf0).max())
    T0 = torch.zeros((N, K), dtype=torch.float32, device="cuda"); T1 = torch.empty_like(T0)
    md = torch.zeros(4, dtype=torch.int32, device="cuda")
    plan.sweep(N, logf, T0, T1, inp, 0.0, md)
    t1 = T1.cpu().numpy(); t1ref = npref.softmax_ref_epilogue(lf0, pk0)
    print(" sweep(beta=0): max abs err", np.abs(t1 - t1ref).max(), "maxdiff", md[:1].view(torch.float32).item(), "expected", np.abs(t1ref).max())
    plan.mstep_accumulate(inp.bits, T1)
    cnt, fsum, nkc, nkf = [t.cpu().numpy() for t in plan.stats_tensors()]
    mu1, eps1, pk1, s1, nk = npref.mstep(sp.X, t1, fd)
    s1g = cnt[:, :D] + fsum[:, :D]; nkg = nkc + nkf
    print(" mstep: S1 max abs err", np.abs(s1g - s1).max(), "nk err", np.abs(nkg - nk).max(), "onehot rows", int(nkc.sum()), "of", N,
          "pad cols nonzero", int(np.abs(cnt[:, D:]).sum() + np.abs(fsum[:, D:]).sum()))
    plan.mstep_finalize()
    mug, epsg, pkg, st = plan.get_params()
    print(" finalize: mu mismatches", int((mug != mu1).sum()), "eps max rel err", float(np.abs(epsg - eps1).max() / eps1.max()),
          "pk err", np.abs(pkg - pk1).max(), "status", st, "halves", int((mu1 == 0.5).sum()))
    plan.density(inp.bits, logf)
    lf = logf.cpu().numpy(); lf1 = npref.logf(sp.X, mug, epsg)
    print(" density(iter1): max abs err", np.abs(lf - lf1).max())
    plan.close()


def whole(N, D, K, beta, fd, itermax=100, seed=42, flags=engine.EPI_REF, host=False, chain=False):
    sp = synth_pangenome(N, D, seed, chain_order=chain)
    be = sp.beta_eff(beta) if beta else 0.0
    t = time.time()
    if host:
        run = engine.nem_run_host(sp.bits, D, sp.rowptr, sp.col, sp.w, K, be, fd, itermax, flags=flags)
        Tg = run.T; labg, entg = engine.labels_device(torch.from_numpy(Tg).cuda())
    else:
        inp = engine.DeviceInput.from_numpy(sp.bits, D, sp.rowptr, sp.col, sp.w)
        run = engine.nem_run_device(inp, K, be, fd, itermax, flags=flags)
        Tg = run.T.cpu().numpy(); labg, entg = engine.labels_device(run.T)
    tg = time.time() - t
    mode = refnem.STRICT if flags == engine.EPI_REF else refnem.FP64
    o = refnem.oracle_nem_run(sp.X, sp.rowptr, sp.col, sp.w, K, be, fd, itermax, mode=mode)
    labo, T3o, ento = refnem.oracle_labels(o["T"])
    labg = labg.cpu().numpy()
    agree = (labg == labo).mean()
    relM = abs(run.crit[3] - o["crit"][3]) / abs(o["crit"][3]) if o["status"] == 0 and run.status == 0 else float("nan")
    print(f"whole N={N} D={D} K={K} beta={be:.3f} fd={fd} flags={flags} host={host} chain={chain}: status g/o {run.status}/{o['status']} "
          f"iters {run.n_iter}/{o['n_iter']} labels agree {agree:.5f} relM {relM:.2e} ent {entg:.4f}/{ento:.4f} "
          f"maxTdiff {np.abs(Tg - o['T']).max():.2e} levels {run.n_levels} em_ms {run.elapsed_ms:.2f} wall {tg:.2f}s")
    print("   crit gpu", ["%g" % c for c in run.crit], "oracle", ["%g" % c for c in o["crit"]])


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "lib abi", engine.lib().nemb200_abi_version())
    for a in [(1500, 200, 3, False), (1000, 70, 5, True), (3000, 1000, 3, False), (2000, 500, 12, True)]:
        try:
            steps(*a)
        except Exception:
            traceback.print_exc()
    for a in [dict(N=2000, D=200, K=3, beta=2.5, fd=False), dict(N=2000, D=200, K=3, beta=2.5, fd=True),
              dict(N=3000, D=1000, K=3, beta=2.5, fd=False), dict(N=4000, D=500, K=5, beta=0, fd=False, itermax=10),
              dict(N=4000, D=500, K=12, beta=0, fd=True, itermax=10), dict(N=3000, D=300, K=2, beta=0, fd=False, itermax=10),
              dict(N=2000, D=200, K=3, beta=2.5, fd=False, host=True), dict(N=2000, D=200, K=3, beta=2.5, fd=False, chain=True),
              dict(N=2000, D=3000, K=3, beta=2.5, fd=False, flags=engine.EPI_LOGSPACE),
              dict(N=3000, D=500, K=20, beta=0, fd=False, itermax=10)]:
        try:
            whole(**a)
        except Exception:
            traceback.print_exc()
