#!/usr/bin/env python3
"""Secondary measurements on the other BASELINE.json configs (parity-test shapes, not the bench headline):
cfg2 single NEM run 20 000 x 1 000; cfg3 K sweep 2..20 on 50 000 families x 500 sampled genomes.
Prints one JSON object per config.  Run on the GPU box: python tools/bench_configs.py [--ref]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome

with_ref = "--ref" in sys.argv


def timed(fn, reps=3):
    best = 1e9
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, out


def cfg2():
    sp = synth_pangenome(20000, 1000, seed=42)
    be = sp.beta_eff(2.5)
    inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    engine.nem_run_device(inp, 3, be)
    dt, run = timed(lambda: engine.nem_run_device(inp, 3, be))
    dth, runh = timed(lambda: engine.nem_run_host(sp.bits, sp.D, sp.rowptr, sp.col, sp.w, 3, be))
    out = {"config": "cfg2: 20000 families x 1000 genomes, K=3, beta=2.5 (eff %.2f), sk_" % be, "iterations": run.n_iter,
           "levels": run.n_levels, "device_resident_s": dt, "em_device_ms": run.elapsed_ms, "host_buffers_s": dth,
           "em_it_per_s_device": run.n_iter / (run.elapsed_ms / 1000), "em_it_per_s_host_call": runh.n_iter / dth}
    if with_ref:
        from oracle import refnem
        t = time.perf_counter()
        rc, ref = refnem.ref_nem_run(sp.X, sp.rowptr, sp.col, sp.w, 3, be)
        tr = time.perf_counter() - t
        lab, _ = engine.labels_device(run.T)
        rlab, _, _ = refnem.oracle_labels(ref["T3"].astype(np.float32))
        out.update({"reference_1core_s_incl_text_io": tr, "reference_iterations": ref["n_iter"],
                    "labels_agree": float((lab.cpu().numpy() == rlab).mean()),
                    "rel_M": abs(run.crit[3] - ref["crit"][3]) / abs(ref["crit"][3])})
    print(json.dumps(out), flush=True)


def cfg3():
    sp = synth_pangenome(50000, 500, seed=43)
    inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
    ks = list(range(2, 21))
    inst = [(inp, k, 0.0, False, 10) for k in ks]
    engine.nem_run_batch(inst[:2])
    tb, runs = timed(lambda: engine.nem_run_batch(inst))
    ts, _ = timed(lambda: [engine.nem_run_device(inp, k, 0.0, False, 10) for k in ks], reps=2)
    its = sum(r.n_iter for r in runs)
    out = {"config": "cfg3: K sweep 2..20 (19 instances, sum K = 209) on 50000 families x 500 genomes, beta 0, <=10 iterations",
           "batched_s": tb, "one_call_per_K_s": ts, "total_iterations": its, "em_it_per_s_batched": its / tb}
    if with_ref:
        from oracle import refnem
        tot = 0.0
        per = {}
        for k in (3, 12):
            t = time.perf_counter()
            rc, ref = refnem.ref_nem_run(sp.X, None if False else sp.rowptr, sp.col, sp.w, k, 0.0, False, 10)
            per[k] = (time.perf_counter() - t, ref["n_iter"] if ref else None)
        out["reference_1core_s_per_K_incl_text_io"] = {str(k): v for k, v in per.items()}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    cfg2()
    cfg3()
