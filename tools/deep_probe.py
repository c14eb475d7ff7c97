#!/usr/bin/env python3
"""Sweep cost on deep Gauss-Seidel schedules (family order that follows the genomes): one from-scratch run of a synthetic
instance, timed; prints the schedule depth, the time and a digest of the posteriors (to compare sweep variants bit for bit).
usage: tools/deep_probe.py <chain_order 0|1|2> <families> <genomes>     (2: a pure chain, one row per level)
env:   NEMB200_SPEC_DEEP=<levels>  NEMB200_SPEC_SWEEP=0"""
import hashlib, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome

mode, N, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sp = synth_pangenome(N, D, seed=7, chain_order=(mode >= 1))
rowptr, col, w = sp.rowptr, sp.col, sp.w
if mode == 2:     # pure chain: row i linked to i - 1 and i + 1
    deg = np.full(N, 2); deg[0] = deg[-1] = 1
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = np.concatenate([[1]] + [[i - 1, i + 1] for i in range(1, N - 1)] + [[N - 2]]).astype(np.int32)
    w = np.full(len(col), 0.5, np.float32)
inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, rowptr, col, w)
be = sp.beta_eff(2.5)
flags = engine.EPI_LOGSPACE if D > 1500 else 0
r = engine.nem_run_device(inp, 3, be, flags=flags, want_params=False)
torch.cuda.synchronize()
t = time.perf_counter()
r = engine.nem_run_device(inp, 3, be, flags=flags, want_params=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t
dig = hashlib.sha1(r.T.cpu().numpy().tobytes()).hexdigest()[:12]
print(f"mode {mode} N {N} D {D} levels {inp.n_levels} iterations {r.n_iter} call_ms {dt*1e3:.3f} em_ms {r.em_ms:.3f} digest {dig} M {r.crit[3]:.6f}")
