#!/usr/bin/env python3
"""One profiled call of a batched workload, bracketed by cudaProfilerStart/Stop (run under
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv`).
usage: tools/profile_batch.py chunk<n>[fd][5]|cfg3|cfg2|cfg5|cfg5fd|shard25k     (chunk16: 16 chunk instances, fd: free dispersion, 5: K = 5)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome, synth_pangenome_large

mode = sys.argv[1]
dev = torch.device("cuda", 0)
if mode.startswith("chunk"):
    fd = "fd" in mode
    kk = 5 if mode.endswith("5") else 3        # chunk16fd5: the K the K-selection picks on the synthetic configs[3] pangenome
    inst = []
    import re
    for b in range(int(re.match(r"chunk(\d+)", mode).group(1))):
        sp = synth_pangenome_large(100000, 500, seed=100 + b, device=dev)
        g = engine.DeviceInput.from_numpy(np.zeros((1, 4), np.uint32), sp.D, sp.rowptr, sp.col, sp.w, device=dev)
        inst.append((engine.DeviceInput(sp.bits, sp.D, g.rowptr, g.col, g.w, g.sched_ptr, g.sched_rows, g.n_levels), kk, sp.beta_eff(2.5), fd, 100))
    call = lambda: engine.nem_run_batch(inst, want_params=False)
elif mode == "cfg3":
    sp = synth_pangenome(50000, 500, seed=43)
    inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    inst = [(inp, k, 0.0, False, 10) for k in range(2, 21)]
    call = lambda: engine.nem_run_batch(inst, want_params=False)
elif mode in ("cfg5", "shard25k", "cfg5fd"):   # shard25k: the rows one of 8 ranks holds of the benchmark instance; cfg5fd: free dispersion
    sp = synth_pangenome_large(25000 if mode == "shard25k" else 200000, 50000, seed=42, device=dev)
    g = engine.DeviceInput.from_numpy(np.zeros((1, 4), np.uint32), sp.D, sp.rowptr, sp.col, sp.w, device=dev)
    inp = engine.DeviceInput(sp.bits, sp.D, g.rowptr, g.col, g.w, g.sched_ptr, g.sched_rows, g.n_levels)
    be = sp.beta_eff(2.5)
    call = lambda: [engine.nem_run_device(inp, 3, be, mode == "cfg5fd", 100, flags=engine.EPI_LOGSPACE, want_params=False)]
else:
    sp = synth_pangenome(20000, 1000, seed=42)
    inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    be = sp.beta_eff(2.5)
    call = lambda: [engine.nem_run_device(inp, 3, be, want_params=False)]
call(); call()
torch.cuda.synchronize()
t = time.perf_counter()
torch.cuda.profiler.start()
runs = call()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
dt = time.perf_counter() - t
print(mode, "call ms", dt * 1e3, "iterations", [r.n_iter for r in runs], "device ms", runs[0].elapsed_ms, "em ms", runs[0].em_ms)
