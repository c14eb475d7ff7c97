#!/usr/bin/env python3
"""Times the density kernel alone on the bench matrix (200 000 x 50 000) for several K, with centres that are not
degenerate (one M-step on random one-hot posteriors first).  Prints ms per launch and GB/s of matrix bytes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome_large

N, D = 200000, 50000
dev = torch.device("cuda", 0)
sp = synth_pangenome_large(N, D, 42, device=dev)
for K in [int(k) for k in (sys.argv[1:] or ["2", "3", "4", "5", "6", "8", "12", "20"])]:
    for fd in (False, True):
        if fd and K not in (3, 8):
            continue
        lab = torch.from_numpy(np.random.default_rng(K).integers(0, K, N)).to(dev)
        T = torch.zeros((N, K), dtype=torch.float32, device=dev)
        T[torch.arange(N, device=dev), lab] = 1.0
        plan = engine.Plan(N, N, D, sp.bits.shape[1], K, fd, engine.EPI_LOGSPACE)
        plan.init_params()
        plan.mstep_accumulate(sp.bits, T); plan.mstep_finalize()
        logf = torch.empty((N, K), dtype=torch.float64, device=dev)
        for _ in range(3):
            plan.density(sp.bits, logf)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        reps = 10
        for _ in range(reps):
            plan.density(sp.bits, logf)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"K={K:2d} {'skd' if fd else 'sk_'}: {ms:.4f} ms/launch, {N * sp.bits.shape[1] * 4 / ms / 1e6:.0f} GB/s", flush=True)
        plan.close()
