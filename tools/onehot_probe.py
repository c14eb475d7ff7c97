#!/usr/bin/env python3
"""Times the one-hot column-count kernel alone on the bench shape (fixed posteriors, full recount each call), for the
kernel variants selected by NEMB200_ONEHOT_RING / NEMB200_OH_DBG.  Prints ms per launch and GB/s."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome_large

N, D, K = 200000, 50000, 3
dev = torch.device("cuda", 0)
sp = synth_pangenome_large(N, D, 42, device=dev)
lab = torch.from_numpy(np.random.default_rng(0).integers(0, K, N)).to(dev)
lab, _ = torch.sort(lab) if os.environ.get("SORTED") else (lab, None)
T = torch.zeros((N, K), dtype=torch.float32, device=dev)
T[torch.arange(N, device=dev), lab] = 1.0
plan = engine.Plan(N, N, D, sp.bits.shape[1], K, False, engine.EPI_LOGSPACE | engine.FULL_RECOUNT)
plan.init_params()
for _ in range(5):
    plan.mstep_accumulate(sp.bits, T)
torch.cuda.synchronize()
plan.timing(True)
for _ in range(30):
    plan.mstep_accumulate(sp.bits, T)
torch.cuda.synchronize()
ms, n = plan.timing(False)
cnt = plan.stats_tensors()[0]
print(f"ring={os.environ.get('NEMB200_ONEHOT_RING')} dbg={os.environ.get('NEMB200_OH_DBG')}: {ms / n:.4f} ms/launch, "
      f"{N * sp.bits.shape[1] * 4 / (ms / n) / 1e6:.0f} GB/s, checksum {int(cnt.to(torch.int64).sum())}")
