#!/usr/bin/env python3
"""Chunk-sample batch (BASELINE configs[3] shape): B independent instances of 100 000 families x 500 genomes, K = 3, beta 2.5,
through nemb200_nem_device_batch, against the same instances one call at a time.
usage: tools/bench_chunk_batch.py [B=8] [free_dispersion=0]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fd = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
insts = []
for b in range(B):
    sp = synth_pangenome(100000, 500, seed=100 + b)
    insts.append((engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w), 3, sp.beta_eff(2.5), fd, 100))
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    runs = engine.nem_run_batch(insts)
    torch.cuda.synchronize(); tb = time.perf_counter() - t
    torch.cuda.synchronize(); t = time.perf_counter()
    single = [engine.nem_run_device(i[0], i[1], i[2], i[3], i[4]) for i in insts]
    torch.cuda.synchronize(); ts = time.perf_counter() - t
its = sum(r.n_iter for r in runs)
same = all(torch.equal(a.T, b.T) for a, b in zip(runs, single))
print(f"B={B} fd={fd}: batched {tb * 1e3:.2f} ms ({its / tb:.0f} EM it/s), one call each {ts * 1e3:.2f} ms, iterations {its}, identical {same}")
