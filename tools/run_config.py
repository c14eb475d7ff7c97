#!/usr/bin/env python3
"""A whole `partition` workflow (K selection -> un-chunked run or chunked votes -> labels) at the shapes BASELINE.json
names, on a synthetic pangenome held as ARRAYS in HBM (no Python gene objects exist at these scales):

    python tools/run_config.py cfg3|cfg4|cfg5|cfg5-chunked [--families F] [--genomes G] [--seed S]
    python -m torch.distributed.run --nproc-per-node N tools/run_config.py cfg4          (chunks / K candidates over ranks)

cfg3          5 000 genomes x 50 000 families: ICL K sweep 2..20 on a 500-genome sample, then chunked partition (chunk 500)
cfg4          20 000 genomes x 100 000 families: chunked subsampling, chunk_size 500, free dispersion (~1 600 chunk instances)
cfg5          50 000 genomes x 200 000 families: K sweep 2..20 on a 500-genome sample, then ONE un-chunked instance with the
              chosen K (row-sharded over the ranks of the process group when there is one)
cfg5-chunked  the same pangenome through the chunked mode (chunk 500: ~10 000 chunk instances)

Synthetic model: families x genomes Bernoulli matrix with the U-shaped frequencies of ppanggolin_b200/synthetic.py;
edges = backbone neighbours at offsets 1 and 2 plus 2 % hub families; an edge has one gene pair in every genome that
holds both families (co-presence model of the exporter).  Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from ppanggolin_b200.nem import engine, workflow
from ppanggolin_b200.nem.export_device import DevicePangenome
from ppanggolin_b200.synthetic import _family_probs, row_stride_words

SHAPES = {"cfg3": (50_000, 5_000), "cfg4": (100_000, 20_000), "cfg5": (200_000, 50_000), "cfg5-chunked": (200_000, 50_000)}


def synth_device_pangenome(F, G, seed, dev):
    rng = np.random.default_rng(seed)
    cls, p = _family_probs(F, rng)
    W = row_stride_words(G)
    gen = torch.Generator(device=dev)
    bits = torch.empty((F, W), dtype=torch.int32, device=dev)
    shifts = torch.arange(32, device=dev, dtype=torch.int64)
    pt = torch.from_numpy(p.astype(np.float32)).to(dev)
    step = 2048
    for r0 in range(0, F, step):
        r1 = min(F, r0 + step)
        gen.manual_seed(seed * 1_000_003 + r0)
        x = torch.rand((r1 - r0, W * 32), device=dev, generator=gen) < pt[r0:r1, None]
        x[:, G:] = False
        x[~x.any(1), 0] = True
        words = (x.view(r1 - r0, W, 32).to(torch.int64) << shifts).sum(-1)
        bits[r0:r1] = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
    backbone = rng.permutation(F)
    e = [np.stack([backbone, np.roll(backbone, -1)], 1), np.stack([backbone, np.roll(backbone, -2)], 1)]
    hubs = rng.choice(F, max(1, F // 50), replace=False)
    for h in hubs:
        other = rng.choice(F, 12, replace=False)
        other = other[other != h]
        e.append(np.stack([np.full(len(other), h), other], 1))
    edges = np.concatenate(e)
    edges = np.unique(np.sort(edges, axis=1), axis=0)
    return DevicePangenome.synthetic_copresence(bits, G, edges.astype(np.int32)), cls


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(SHAPES))
    ap.add_argument("--families", type=int, default=None)
    ap.add_argument("--genomes", type=int, default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--chunk-size", type=int, default=500)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    F, G = SHAPES[a.config]
    F = a.families or F; G = a.genomes or G
    t0 = time.perf_counter()
    dp, cls = synth_device_pangenome(F, G, a.seed, dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    random.seed(1234)
    fd = a.config == "cfg4"
    out = {"config": a.config, "families": F, "genomes": G, "edges": dp.E, "n_gpus": world, "free_dispersion": fd,
           "chunk_size": a.chunk_size, "generate_s": t_gen}
    t0 = time.perf_counter()
    if a.config == "cfg5":
        # K from the sweep on a 500-genome sample (partition.py:474-477 with the default chunk size), then one instance over
        # all genomes with that K (partition(..., chunk_size >= number of genomes))
        k, bics, icls, lls = workflow.evaluate_nb_partitions_arrays(dp, list(range(G)), fd, a.chunk_size, [3, 20], 0.05)
        torch.cuda.synchronize()
        out["k_selection_s"] = time.perf_counter() - t0
        out["icl"] = {int(kk): float(v) for kk, v in icls.items()}
        res = workflow.partition_arrays(dp, 2.5, 10, fd, chunk_size=G, kval=k, seed=42, want_params=False)
    else:
        res = workflow.partition_arrays(dp, 2.5, 10, fd, chunk_size=a.chunk_size, kval=-1, krange=[3, 20], seed=42, want_params=False)
    torch.cuda.synchronize()
    out["partition_wall_s"] = time.perf_counter() - t0
    out["K"] = res.K
    out["chunk_instances"] = len(res.samples)
    out["timings"] = {k: (float(v) if not isinstance(v, int) else v) for k, v in res.timings.items()}
    if res.timings.get("em_s") and res.timings.get("iterations"):
        out["em_it_per_s_this_rank"] = res.timings["iterations"] / res.timings["em_s"]
    letters = np.array([str(l)[:1] for l in res.labels])
    truth = np.array(["P", "S", "C"])[cls]
    out["label_counts"] = {l: int((letters == l).sum()) for l in ("P", "S", "C", "U", "")}
    out["letter_agreement_with_generator_classes"] = float((letters == truth).mean())
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
