#!/usr/bin/env python3
"""Benchmark of the NEM partitioning hot path (BASELINE.json metric: NEM EM iterations/sec vs the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--families F] [--genomes G]

Workload (config.workload): BASELINE.json configs[4] -- synthetic 50 000 genomes x 200 000 families, one un-chunked NEM
instance (K = 3, beta = 2.5, dispersion "sk_"), row-sharded over the N GPUs (strong scaling: the pangenome is fixed).
It is the configuration the metric is quoted on ("at 1/2/4/8 B200") and fits one GPU (1.25 GB of bits, 10x the L2: no
flush needed between steps).

A *step* is one EM iteration of a run that STARTS FROM THE INITIAL PARAMETERS (partition.py:65-85) and runs to
convergence, exactly what `run_partitioning` does: start phase (initial densities, blind sweep, sweep with beta), then
per iteration M-step -> densities -> Gauss-Seidel posterior sweep -> convergence test.  The timed region is a sequence
of such runs, back to back, until exactly K iterations have been executed (the last run is cut by its iteration cap); the
start phase, the per-run set-up and the final criteria of every run are inside the region.

`value`   : K / (CUDA-event time of that region), product defaults (incremental M-step: rows that did not change class are
            not recounted -- integer arithmetic, bit-identical to a recount).  N = 1: through the C ABI call
            `nemb200_nem_device` (device-resident loop, no host synchronisation inside the EM); N > 1: through the
            row-sharded driver (nem/sharded.py, exchange over NVLink peer memory), max over ranks.
`value_full_recount`: the same with every one-hot row recounted at every iteration (each step streams the matrix twice).
`e2e`     : the same metric through the host-buffer C-ABI call `nemb200_nem_host` (pinned host matrix -> H2D -> EM to
            convergence -> posteriors D2H), i.e. what a caller of `run_partitioning` pays per instance.
`roofline`: the dominant kernel of an iteration (algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json) in
            `achieved`/`frac`; `iteration_frac` is SURVEY.md section 8(d)'s figure B_iter / t_iter / peak over the timed
            runs (matrix counted once per iteration).
`configs` : (N = 1) the other BASELINE configs through the batched device-resident path -- configs[1] one run,
            configs[2] the 19-candidate K sweep, configs[3] a batch of 16 chunk instances (free dispersion) -- each with
            the UNMODIFIED reference C timed on the same instance (or a stated row sample of it) on the host cores.
`cpu_baseline` / `--impl reference`: at this width (50 000 genomes) the reference underflows and stops after one
            degenerate iteration, so the CPU leg is the fp64 restatement (oracle/nem_oracle.c,
            kind "port") on row samples of the same matrix, one process per host core, extrapolated linearly in rows.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K_CLASSES = 3
BETA = 2.5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--families", type=int, default=200_000)
    ap.add_argument("--genomes", type=int, default=50_000)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-sample-rows", type=int, default=160)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--nccl", action="store_true", help="N>1: exchange through NCCL collectives instead of the kernels' NVLink peer-memory path")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region: NVML in-process (10 ms period, no subprocesses that could
    disturb the run); falls back to polling nvidia-smi."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.sm, self.reason_bits, self.max_mhz, self._halt = index, [], 0, None, threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self.sm.append(int(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM)))
                    try:
                        self.reason_bits |= int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        self.reason_bits |= int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self._halt.wait(0.01)
                else:
                    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(int(out[0])); self.max_mhz = int(out[1]); self.reason_bits |= int(out[2].strip(), 16)
                    self._halt.wait(0.2)
            except Exception:
                self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        sm = sorted(self.sm)
        reasons = [n for n, bit in self.REASONS.items() if self.reason_bits & bit]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def b_iter_bytes(N, D, K, nnz):
    """algorithmic bytes of one EM iteration (matrix counted once)."""
    return N * ((D + 7) // 8) + 2 * N * K * 4 + (N + 1) * 4 + nnz * 8 + nnz * K * 4 + 3 * K * D * 4


# ------------------------------------------------------------------------------------------------------- CPU legs
def _port_rows_worker(args):
    """One host process: the fp64 restatement (oracle/nem_oracle.c) on a row sample of the wide matrix."""
    X, rowptr, col, w, K, beta = args
    from oracle import refnem
    t = time.time()
    o = refnem.oracle_nem_run(X, rowptr, col, w, K, beta, False, 100, mode=refnem.FP64)
    return max(o["n_iter"], 1), time.time() - t, o["status"]


def _chain_graph(pc):
    """Chain neighbourhood of a row sample (each family linked to the next one), weights ~ joint presence."""
    rows = len(pc)
    col, w, rowptr = [], [], [0]
    for i in range(rows):
        for j in (i - 1, i + 1):
            if 0 <= j < rows:
                col.append(j); w.append(round(float(pc[i] * pc[j]), 4) or 0.0001)
        rowptr.append(len(col))
    return np.array(rowptr, np.int32), np.array(col, np.int32), np.array(w, np.float32)


def cpu_wide_sample(N_full, D, rows, seed, K, beta, cores):
    """EM it/s of the fp64 restatement on `rows` sampled families x D genomes per core, scaled linearly to N_full rows."""
    from multiprocessing import get_context
    from oracle import refnem
    from ppanggolin_b200.synthetic import _family_probs
    refnem.build(ref=False)
    rng = np.random.default_rng(seed + 1)
    cls, p = _family_probs(rows * cores, rng)
    jobs = []
    for c in range(cores):
        pc = p[c * rows:(c + 1) * rows]
        X = (rng.random((rows, D), dtype=np.float32) < pc[:, None].astype(np.float32)).astype(np.uint8)
        X[X.sum(1) == 0, 0] = 1
        rowptr, col, w = _chain_graph(pc)
        jobs.append((X, rowptr, col, w, K, beta * rows / (float(w.sum()) / 2)))
    t0 = time.time()
    with get_context("fork").Pool(cores) as pool:
        res = pool.map(_port_rows_worker, jobs)
    wall = time.time() - t0
    its = sum(n / dt for n, dt, _ in res)                 # sample-size iterations / s over all cores
    return {"value": its * rows / N_full, "unit": "EM iterations/s", "cores": cores, "kind": "port",
            "sample": f"fp64 restatement (oracle/nem_oracle.c; the reference C underflows at {D} genomes and stops after one "
                      f"degenerate iteration): {cores} processes x ({rows} families x {D} genomes, K={K}, run to convergence), "
                      f"scaled linearly to {N_full} families; iterations={[r[0] for r in res]}", "wall_s": wall}


def _ref_job(args):
    """One host process: the UNMODIFIED reference C (oracle/_ref) or, where it was not built, the STRICT restatement, on an
    instance given as arrays; returns (tag, iterations, seconds of nem() only, labels | None)."""
    tag, X, rowptr, col, w, K, beta, fd, itermax, want_labels = args
    from oracle import refnem
    if refnem.have_ref():
        import shutil
        import tempfile
        d = Path(tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp")))
        refnem.write_nem_files_fast(d, X, rowptr, col, w)
        refnem.write_init_file(d, K, X.shape[1])
        t = time.time()
        refnem.ref_nem_call(d, K, float(np.float32(beta)), fd, itermax)
        dt = time.time() - t
        out = refnem.parse_ref_outputs(d, K, X.shape[1])
        shutil.rmtree(d, ignore_errors=True)
        if out is None:
            return tag, 0, dt, None, None, "reference"
        lab = refnem.oracle_labels(out["T3"].astype(np.float32))[0] if want_labels else None
        return tag, out["n_iter"] or itermax, dt, lab, out["crit"][3], "reference"
    t = time.time()
    o = refnem.oracle_nem_run(X, rowptr, col, w, K, beta, fd, itermax, mode=refnem.STRICT)
    dt = time.time() - t
    lab = refnem.oracle_labels(o["T"])[0] if want_labels else None
    return tag, o["n_iter"], dt, lab, o["crit"][3], "port"


def _restrict_rows(sp_X, rowptr, col, w, n):
    """First n rows of an instance, neighbour lists cut to those rows."""
    rp = [0]; cc = []; ww = []
    for i in range(n):
        a, b = int(rowptr[i]), int(rowptr[i + 1])
        keep = col[a:b] < n
        cc.append(col[a:b][keep]); ww.append(w[a:b][keep]); rp.append(rp[-1] + int(keep.sum()))
    return sp_X[:n], np.array(rp, np.int32), np.concatenate(cc).astype(np.int32), np.concatenate(ww).astype(np.float32)


# ---------------------------------------------------------------------------------------------------- reference arm
def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 32)
    vals = []
    t0 = time.time()
    budget_s = 200.0
    for s in range(a.warmup + a.steps):
        r = cpu_wide_sample(a.families, a.genomes, a.cpu_sample_rows, a.seed + s, K_CLASSES, BETA, cores)
        if s >= a.warmup:
            vals.append(r)
        if time.time() - t0 > budget_s and vals:
            break
    v = float(np.mean([r["value"] for r in vals]))
    line = {"metric": "NEM EM iterations/sec", "value": v, "unit": "EM iterations/s", "impl": "reference",
            "n_gpus": a.gpus, "steps": len(vals), "warmup": a.warmup, "ms_per_step": 1000.0 / v if v > 0 else None,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 (CPU restatement of the reference)",
            "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": v, "unit": "EM iterations/s", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"],
                             "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "EM iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(a):
    return {"workload": f"cfg5: synthetic {a.genomes} genomes x {a.families} families, one un-chunked NEM instance run from the "
                        f"initial parameters to convergence, K={K_CLASSES}, beta={BETA}, dispersion sk_, Gauss-Seidel sweep, "
                        f"log-space epilogue, row-sharded over the GPUs", "families": a.families, "genomes": a.genomes,
            "K": K_CLASSES, "beta": BETA, "l2_policy": "matrix (families x genomes bits) is larger than L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------------- our arm
def timed_runs(run_once, steps, barrier, torch, st):
    """Back-to-back from-scratch runs until exactly `steps` EM iterations were executed.  `run_once(itermax)` -> (n_iter,
    em_ms, run_ms, launches).  Returns (CUDA-event ms of the region, runs, iterations per run, sum em_ms, sum run_ms)."""
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    done, per_run, em_ms, run_ms, launches = 0, [], 0.0, 0.0, 0
    barrier()
    t0.record(st)
    while done < steps:
        n, e, r, l = run_once(min(100, steps - done))
        n = max(n, 1)
        done += n; per_run.append(n); em_ms += e; run_ms += r; launches += l
    t1.record(st)
    barrier()
    return t0.elapsed_time(t1), per_run, em_ms, run_ms, launches


def launches_of_run(n_iter, k_launch_per_iter, start, end):
    return start + n_iter * k_launch_per_iter + end


def run_ours(a):
    import torch
    import torch.distributed as dist
    from ppanggolin_b200.nem import engine
    from ppanggolin_b200.nem.sharded import CudaBackend, ShardedEM, shard_rows
    from ppanggolin_b200.synthetic import synth_pangenome_large

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    N, D, K = a.families, a.genomes, K_CLASSES
    lo, hi = shard_rows(N, world, rank)

    # synthetic pangenome: every rank generates the same full graph (tiny) and its own rows of the matrix
    sp = synth_pangenome_large(N, D, a.seed, device=dev, rows=(lo, hi))
    beta = float(np.float32(sp.beta_eff(BETA)))
    graph = engine.DeviceInput.from_numpy(np.zeros((1, 4), np.uint32), D, sp.rowptr, sp.col, sp.w, device=dev)
    nnz = int(len(sp.col))
    st = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=group)
            torch.cuda.synchronize()

    pg = group if (world > 1 and not a.nccl) else None
    last = {}
    dens = {"ms": 0.0, "n": 0}     # density launches of the timed runs: device time (library events around each launch) and count
    if world == 1:
        inp = engine.DeviceInput(sp.bits, D, graph.rowptr, graph.col, graph.w, graph.sched_ptr, graph.sched_rows, graph.n_levels)

        def make_run(flags):
            def run_once(itermax):
                r = engine.nem_run_device(inp, K, beta, False, itermax, 0.01, flags, want_params=False)
                last["run"] = r
                dens["ms"] += r.density_ms; dens["n"] += r.density_launches
                # kernels of one run of the device-resident loop (csrc/nemb200_batch.inc): start (start kernel, graph copy,
                # densities, 2 sweeps), per iteration one-hot counts / fuzzy sums / finalize / densities / sweep (its tail
                # carries the convergence decision and the next classification), one gated iteration in flight when
                # convergence is seen, then criteria, their reduction and collect
                return r.n_iter, r.em_ms, r.elapsed_ms, 5 + 5 * (r.n_iter + 1) + 3
            return run_once
        be = None
    else:
        be = CudaBackend(sp.bits, D, N, K, False, engine.EPI_LOGSPACE | engine.TIME_DENSITY, graph, peer_group=pg, row_offset=lo)
        be_full = CudaBackend(sp.bits, D, N, K, False, engine.EPI_LOGSPACE | engine.FULL_RECOUNT, graph, peer_group=pg, row_offset=lo)

        def make_run(flags):
            b = be_full if (flags & engine.FULL_RECOUNT) else be

            def run_once(itermax):
                b.launches = 0
                b.density_ms, b.density_launches = 0.0, 0
                out = ShardedEM(b, N, K, rank, world, group).run(beta, itermax, 0.01)
                last["out"] = out
                dens["ms"] += b.density_ms; dens["n"] += b.density_launches
                return out.n_iter, 0.0, 0.0, b.launches
            return run_once

    run_default = make_run(engine.EPI_LOGSPACE | engine.TIME_DENSITY)
    run_full = make_run(engine.EPI_LOGSPACE | engine.FULL_RECOUNT)
    # warm-up: whole runs until at least a.warmup iterations were executed (allocator cache, module load, clocks)
    w_done = 0
    while w_done < max(a.warmup, 3):
        w_done += max(run_default(100)[0], 1)
    run_full(100)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    dens["ms"], dens["n"] = 0.0, 0
    profile_region = os.environ.get("NEMB200_BENCH_PROFILE") is not None   # ncu --profile-from-start off: only the timed region
    if profile_region:
        torch.cuda.profiler.start()
    ms_total, per_run, em_ms, run_ms, launches = timed_runs(run_default, a.steps, barrier, torch, st)
    if profile_region:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None
    dens_ms, dens_n = dens["ms"], dens["n"]
    ms_full, per_run_full, em_ms_full, _, _ = timed_runs(run_full, a.steps, barrier, torch, st)
    if world > 1:
        t = torch.tensor([ms_total, ms_full], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        ms_total, ms_full = float(t[0].item()), float(t[1].item())
    value = a.steps / (ms_total / 1000.0)

    # ---- per-stage / per-kernel times of full-work iterations (stepwise interface, CUDA events between the stages; kept
    # out of the timed region: an event record costs stream time too)
    bp = CudaBackend(sp.bits, D, N, K, False, engine.EPI_LOGSPACE | engine.FULL_RECOUNT, graph, peer_group=pg, row_offset=lo) \
        if world == 1 else be_full
    em = ShardedEM(bp, N, K, rank, world, group)
    em.start(beta)
    for _ in range(3):
        em.iterate(beta)
    n_prof = 12
    bp.plan.timing(True)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(n_prof)]
    for s in range(n_prof):
        e = ev[s]
        e[0].record(st)
        bp.accumulate(bp.T[em.cur][em.lo:em.hi]); e[1].record(st)
        em._allreduce_stats(); bp.finalize(); e[2].record(st)
        bp.density(bp.logf_full[em.lo:em.hi]); e[3].record(st)
        em._allgather_logf(); bp.sweep(bp.T[em.cur], bp.T[1 - em.cur], beta); em.cur = 1 - em.cur
        e[4].record(st)
    barrier()
    onehot_ms, onehot_n = bp.plan.timing(False)
    stage = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in ev])  # accumulate, finalize, density, sweep
    md, status = bp.read_flags()

    # correctness guard inside the bench: labels of the last timed run must recover the generator's classes
    Tf = last["run"].T if world == 1 else last["out"].T
    lab, _ = engine.labels_device(Tf)
    truth = torch.from_numpy(sp.true_class.astype(np.int32)).to(dev)
    acc = float((lab == truth).float().mean().item())
    t_checksum = float((Tf.double() * torch.arange(1, K + 1, device=dev, dtype=torch.float64)).sum().item())
    n_conv = per_run[0]

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        st_mean = stage.mean(0)
        names = ["mstep_accumulate (classify/scatter + k_mstep_onehot_ring with k_mstep_fuzzy beside it)", "mstep_finalize (+ all-reduce)",
                 "density (k_density_sk_ring)", "sweep (k_sweep, level-scheduled; + all-gather)"]
        n_loc = hi - lo
        x_bytes = n_loc * ((D + 7) // 8)
        traffic, traffic_src = {}, None
        for tname in ("r02_traffic.json", "r01_traffic.json"):
            tp = ROOT / "profiles" / tname
            if tp.is_file():
                traffic = json.loads(tp.read_text()); traffic_src = f"profiles/{tname} (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch; not measured in this run)"
                break
        dens_bytes = x_bytes + n_loc * K * 8 + 2 * K * ((D + 7) // 8)
        kernels = {
            "k_b_density_ring": {"ms": dens_ms / max(dens_n, 1), "launches_timed": dens_n, "bytes": dens_bytes,
                                 "what": "matrix bits + log-densities out + centre planes; mean over the density launches of the "
                                         "timed region's EM iterations, library-side CUDA events around each launch on its stream "
                                         "(NEMB200_TIME_DENSITY)" + ("" if world == 1 else "; includes the push of the shard to the peers (kernel epilogue)")},
            "k_density_sk_ring (stepwise entry, same device body, timed alone after the region)": {
                "ms": float(st_mean[2]), "bytes": dens_bytes, "what": "matrix bits + log-densities out + centre planes"},
            "k_mstep_onehot_ring": {"ms": onehot_ms / max(onehot_n, 1), "bytes": x_bytes + n_loc * 4 + K * D * 4,
                                    "what": "matrix bits + row lists + column counts out (full recount)"}}
        for kk, v in kernels.items():
            v["achieved_GBs"] = v["bytes"] / (v["ms"] / 1000.0) / 1e9 if v["ms"] > 0 else None
            v["frac"] = v["achieved_GBs"] / peak if v["ms"] > 0 else None
        dom = "k_b_density_ring"       # runs once in every iteration of every mode; the count kernel only streams changed rows
        biter = b_iter_bytes(N, D, K, nnz)
        tr = traffic.get(dom, {}).get("dram_bytes_per_launch") if world == 1 else None
        t_iter = ms_total / a.steps / 1000.0
        line = {"metric": "NEM EM iterations/sec", "value": value, "unit": "EM iterations/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32 bit ops + f64 epilogue", "data": "synthetic",
                "config": workload_config(a), "gpu_launches": launches, "clocks": clocks,
                "timed_region": {"what": "from-scratch runs to convergence, back to back, until `steps` iterations", "runs": len(per_run),
                                 "iterations_per_run": per_run, "iterations_to_convergence": n_conv,
                                 "em_loop_ms_sum (library events: start phase + iterations)": em_ms if world == 1 else None,
                                 "run_ms_sum (library events: + criteria)": run_ms if world == 1 else None,
                                 "region_ms (CUDA events around the calls: + per-run set-up and result copies)": ms_total},
                "mstep_mode": "incremental (product default)",
                "value_full_recount": {"value": a.steps / (ms_full / 1000.0), "unit": "EM iterations/s", "ms_per_step": ms_full / a.steps,
                                       "iterations_per_run": per_run_full,
                                       "what": "the same runs with every one-hot row recounted at every iteration"},
                "exchange": ("none (1 GPU)" if world == 1 else ("NCCL all-reduce + all-gather" if a.nccl else "NVLink peer memory inside the kernels: statistics / parameters as {payload, epoch} words in k_b_finalize_ll (no fence, no flag), log-densities pushed by the epilogue of k_b_density_ring")),
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBs"], "peak": peak,
                             "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": tr,
                             "traffic_source": traffic_src if tr else None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": kernels[dom]["bytes"], "kernel_ms": kernels[dom]["ms"],
                             "iteration_frac": biter / t_iter / 1e9 / (peak * world),
                             "iteration": {"B_iter_bytes": biter, "t_iter_ms": t_iter * 1e3, "achieved_GBs": biter / t_iter / 1e9,
                                           "frac": biter / t_iter / 1e9 / (peak * world),
                                           "what": "SURVEY 8(d): B_iter / (region time / iterations) / (peak x GPUs), from-scratch runs"},
                             "iteration_full_recount_frac": biter / (ms_full / a.steps / 1000.0) / 1e9 / (peak * world),
                             "kernels": kernels, "stage_ms_full_work_iteration": {n: float(v) for n, v in zip(names, st_mean)}},
                "check": {"label_accuracy_vs_generator": acc, "gs_levels": graph.n_levels, "nnz": nnz,
                          "last_maxdiff": md, "status": status, "posterior_checksum": t_checksum}}
    if world == 1:
        del bp, em
    # ---- N > 1: the instance-parallel mode: independent instances dealt over the ranks, no communication
    if world > 1:
        ip = instance_parallel(engine, torch, dist, group, dev, rank, world)
        if rank == 0:
            line["instance_parallel"] = ip
    # ---- end to end through the host-buffer C ABI
    if not a.no_e2e and world == 1:
        host_bits = torch.empty(sp.bits.shape, dtype=torch.int32, pin_memory=True)
        host_bits.copy_(sp.bits)
        torch.cuda.synchronize()
        hb = host_bits.numpy().view(np.uint32)
        del inp
        last.clear()
        sp.bits = None
        torch.cuda.empty_cache()
        # one untimed call first (device allocator cache, module load), then the timed one: every timed call still moves
        # the whole matrix host -> device and the posteriors back
        engine.nem_run_host(hb, D, sp.rowptr, sp.col, sp.w, K, beta, False, 100, 0.01, engine.EPI_LOGSPACE)
        t0 = time.perf_counter()
        run = engine.nem_run_host(hb, D, sp.rowptr, sp.col, sp.w, K, beta, False, 100, 0.01, engine.EPI_LOGSPACE)
        dt = time.perf_counter() - t0
        its = max(run.n_iter, 1)
        line["e2e"] = {"value": its / dt, "unit": "EM iterations/s", "h2d_bytes_per_step": int(hb.nbytes + sp.col.nbytes * 2 + sp.rowptr.nbytes) // its,
                       "d2h_bytes_per_step": int(N * K * 4 + 2 * K * D * 4) // its, "iterations": its, "wall_s": dt,
                       "em_device_ms": run.elapsed_ms, "converged": run.converged,
                       "call": "nemb200_nem_host (pinned host matrix -> H2D -> EM to convergence -> posteriors D2H)"}
        del host_bits, hb
    elif not a.no_e2e:
        # N > 1: the same job through the row-sharded driver -- every rank's shard starts in pinned host memory, is copied
        # to its GPU, EM runs to convergence over the same peer-memory exchange as `value`, rank 0 reads the posteriors back
        host_bits = torch.empty(sp.bits.shape, dtype=torch.int32, pin_memory=True)
        host_bits.copy_(sp.bits)
        dbits = sp.bits

        def one_job():
            dbits.copy_(host_bits, non_blocking=True)
            res = ShardedEM(be, N, K, rank, world, group).run(beta, 100, 0.01)
            if rank == 0:
                _ = res.T.cpu()
            return res

        one_job()
        barrier()
        t0 = time.perf_counter()
        out = one_job()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        if rank == 0:
            its = max(out.n_iter, 1)
            line["e2e"] = {"value": its / float(t.item()), "unit": "EM iterations/s",
                           "h2d_bytes_per_step": int(host_bits.numel() * 4 * world) // its,
                           "d2h_bytes_per_step": int(N * K * 4) // its, "iterations": its, "wall_s": float(t.item()),
                           "converged": out.converged,
                           "call": "ShardedEM.run (pinned host shards -> H2D -> row-sharded EM to convergence, same exchange as `value` -> posteriors D2H)"}
    elif rank == 0:
        line["e2e"] = None
    if rank == 0 and world == 1 and not a.no_configs:
        try:
            line["configs"] = other_configs(engine, torch, dev, not a.no_cpu)
        except Exception as ex:  # secondary numbers must not take the headline down with them
            line["configs"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not a.no_cpu:   # the CPU baseline is reported at N = 1 only
        cores = min(os.cpu_count() or 1, 32)
        try:
            cb = cpu_wide_sample(N, D, a.cpu_sample_rows, a.seed, K, BETA, cores)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline must not take the GPU numbers down with it
            line["cpu_baseline"] = {"value": None, "unit": "EM iterations/s", "cores": cores, "kind": "unavailable", "sample": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=group)
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------- other BASELINE configs (N = 1)
def _time_best(fn, torch, reps=3):
    best, out = 1e9, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, out


def other_configs(engine, torch, dev, with_cpu: bool):
    """configs[1..3] through the batched device-resident path; the reference timed on the same instances on the host."""
    from ppanggolin_b200.synthetic import synth_pangenome, synth_pangenome_large, unpack_rows
    peak, _ = measured_peaks()
    out = {}
    # configs[1]: one run
    sp2 = synth_pangenome(20000, 1000, seed=42)
    be2 = sp2.beta_eff(2.5)
    in2 = engine.DeviceInput.from_numpy(sp2.bits, sp2.D, sp2.rowptr, sp2.col, sp2.w, device=dev)
    engine.nem_run_device(in2, 3, be2)
    dt2, r2 = _time_best(lambda: engine.nem_run_device(in2, 3, be2), torch)
    b2 = b_iter_bytes(sp2.N, sp2.D, 3, len(sp2.col))
    out["cfg2"] = {"what": "configs[1]: 20000 families x 1000 genomes, K=3, beta 2.5, sk_, one run to convergence (nemb200_nem_device)",
                   "iterations": r2.n_iter, "gs_levels": r2.n_levels, "call_ms": dt2 * 1e3, "device_ms": r2.elapsed_ms, "em_loop_ms": r2.em_ms,
                   "em_it_per_s": r2.n_iter / dt2, "B_iter_bytes": b2, "roofline_frac_B_iter": b2 * r2.n_iter / dt2 / 1e9 / peak}
    # configs[2]: the K sweep as one batch
    sp3 = synth_pangenome(50000, 500, seed=43)
    in3 = engine.DeviceInput.from_numpy(sp3.bits, sp3.D, sp3.rowptr, sp3.col, sp3.w, device=dev)
    ks = list(range(2, 21))
    inst3 = [(in3, k, 0.0, False, 10) for k in ks]
    engine.nem_run_batch(inst3[:2])
    dt3, r3 = _time_best(lambda: engine.nem_run_batch(inst3, want_params=False), torch)
    it3 = sum(r.n_iter for r in r3)
    b3 = sum(b_iter_bytes(sp3.N, sp3.D, k, 0) * r.n_iter for k, r in zip(ks, r3))
    out["cfg3"] = {"what": "configs[2] K selection: 19 candidates K=2..20 on 50000 families x 500 sampled genomes, beta 0, <= 10 iterations, ONE batched call (nemb200_nem_device_batch)",
                   "iterations_total": it3, "call_ms": dt3 * 1e3, "em_it_per_s": it3 / dt3, "B_iter_bytes_total": b3,
                   "roofline_frac_B_iter": b3 / dt3 / 1e9 / peak}
    # configs[3]: a batch of 16 chunk instances, free dispersion
    sps = [synth_pangenome_large(100000, 500, seed=100 + b, device=dev) for b in range(16)]
    inst4 = []
    for sp in sps:
        g = engine.DeviceInput.from_numpy(np.zeros((1, 4), np.uint32), sp.D, sp.rowptr, sp.col, sp.w, device=dev)
        inst4.append((engine.DeviceInput(sp.bits, sp.D, g.rowptr, g.col, g.w, g.sched_ptr, g.sched_rows, g.n_levels), 3,
                      sp.beta_eff(2.5), True, 100))
    engine.nem_run_batch(inst4[:2])
    dt4, r4 = _time_best(lambda: engine.nem_run_batch(inst4, want_params=False), torch)
    it4 = sum(r.n_iter for r in r4)
    b4 = sum(b_iter_bytes(sp.N, sp.D, 3, len(sp.col)) * r.n_iter for sp, r in zip(sps, r4))
    out["cfg4"] = {"what": "configs[3] chunk instances: 16 x (100000 families x 500 genomes), K=3, beta 2.5, free dispersion (skd), ONE batched call",
                   "iterations_total": it4, "iterations": [r.n_iter for r in r4], "call_ms": dt4 * 1e3, "em_it_per_s": it4 / dt4,
                   "B_iter_bytes_total": b4, "roofline_frac_B_iter": b4 / dt4 / 1e9 / peak,
                   "per_instance_ms": dt4 * 1e3 / 16}
    if with_cpu:
        from multiprocessing import get_context
        from oracle import refnem
        refnem.build(ref=False)
        sp4 = sps[0]
        X4 = unpack_rows(sp4.bits.cpu().numpy().view(np.uint32), sp4.D)
        n4 = 25000
        X4s, rp4, c4, w4 = _restrict_rows(X4, sp4.rowptr, sp4.col, sp4.w, n4)
        ew4 = float(w4.sum()) / 2
        jobs = [("cfg2_full", sp2.X, sp2.rowptr, sp2.col, sp2.w, 3, be2, False, 100, True),
                ("cfg2_it1", sp2.X, sp2.rowptr, sp2.col, sp2.w, 3, be2, False, 1, False),
                ("cfg3_K3_full", sp3.X, sp3.rowptr, sp3.col, sp3.w, 3, 0.0, False, 10, False),
                ("cfg3_K3_it1", sp3.X, sp3.rowptr, sp3.col, sp3.w, 3, 0.0, False, 1, False),
                ("cfg4_full", X4s, rp4, c4, w4, 3, 2.5 * n4 / ew4, True, 100, False),
                ("cfg4_it1", X4s, rp4, c4, w4, 3, 2.5 * n4 / ew4, True, 1, False)]
        cores = min(os.cpu_count() or 1, len(jobs))
        with get_context("fork").Pool(cores) as pool:
            res = {r[0]: r for r in pool.map(_ref_job, jobs)}

        def leg(name):
            _, n, t_full, lab, M, kind = res[name + "_full"]
            _, n1, t_1, _, _, _ = res[name + "_it1"]
            per_it = (t_full - t_1) / max(n - n1, 1)
            return {"kind": kind, "cores": 1, "iterations": n, "nem_s_total": t_full, "nem_s_one_iteration_run": t_1,
                    "s_per_iteration": per_it, "em_it_per_s": 1.0 / per_it if per_it > 0 else None}, lab, M
        l2, lab2, M2 = leg("cfg2")
        if lab2 is not None:
            labg, _ = engine.labels_device(r2.T)
            l2["labels_identical_to_gpu_run"] = float((labg.cpu().numpy() == lab2).mean())
            l2["rel_M_vs_gpu_run"] = abs(r2.crit[3] - M2) / abs(M2)
        l2["what"] = "the reference's nem() on the same instance, one core (it is single-threaded); text parse + per-column qsort separated out with a 1-iteration run"
        out["cfg2"]["cpu_reference"] = l2
        l3, _, _ = leg("cfg3_K3")
        l3["what"] = "the reference's nem() for ONE candidate (K=3) of the sweep on the same matrix, one core; the sweep is 19 such processes (K=20 costs ~10x K=3)"
        out["cfg3"]["cpu_reference"] = l3
        l4, _, _ = leg("cfg4")
        l4["what"] = f"the reference's nem() on the first {n4} families of one chunk instance (same columns, graph cut to those rows), one core; a full instance is {100000 // n4}x that"
        out["cfg4"]["cpu_reference"] = l4
    return out


def instance_parallel(engine, torch, dist, group, dev, rank, world):
    """Independent instances dealt over the ranks (no communication during EM): the K-selection sweep of configs[2]
    (19 candidates, round-robin) and chunk batches of configs[3] (16 instances per rank)."""
    from ppanggolin_b200.synthetic import synth_pangenome_large
    sp3 = synth_pangenome_large(50000, 500, seed=43, device=dev)
    in3 = engine.DeviceInput(sp3.bits, sp3.D)
    ks = list(range(2, 21))[rank::world]
    inst3 = [(in3, k, 0.0, False, 10) for k in ks]
    sps = [synth_pangenome_large(100000, 500, seed=1000 + 16 * rank + b, device=dev) for b in range(16)]
    inst4 = []
    for sp in sps:
        g = engine.DeviceInput.from_numpy(np.zeros((1, 4), np.uint32), sp.D, sp.rowptr, sp.col, sp.w, device=dev)
        inst4.append((engine.DeviceInput(sp.bits, sp.D, g.rowptr, g.col, g.w, g.sched_ptr, g.sched_rows, g.n_levels), 3,
                      sp.beta_eff(2.5), True, 100))
    res = {}
    for name, inst in (("k_sweep", inst3), ("chunk_batches", inst4)):
        engine.nem_run_batch(inst[:1], want_params=False)
        best, its = 1e9, 0
        for _ in range(3):
            torch.cuda.synchronize(); dist.barrier(group=group)
            t = time.perf_counter()
            runs = engine.nem_run_batch(inst, want_params=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
            best = min(best, float(tt.item()))
            its = sum(r.n_iter for r in runs)
        ti = torch.tensor([its], dtype=torch.float64, device=dev)
        dist.all_reduce(ti, op=dist.ReduceOp.SUM, group=group)
        res[name] = {"instances_total": (19 if name == "k_sweep" else 16 * world), "ms_max_over_ranks": best * 1e3,
                     "iterations_total": int(ti.item()), "em_it_per_s_all_ranks": float(ti.item()) / best,
                     "scaling": "strong (19 candidates over the ranks)" if name == "k_sweep" else "weak (16 chunk instances per rank)"}
    return res


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
