"""The partitioning workflow of ``partition()`` (/root/reference/ppanggolin/nem/partition.py:653-899) on the array form of a
pangenome: K selection (:439-545), then one un-chunked NEM run (:866-878) or the chunked mode with majority votes
(:773-864).  ``partition.partition`` (Python objects in, ``fam.partition`` out) calls it after one traversal of the
objects; a caller that already holds arrays -- a pangenome of 10^4..10^5 genomes cannot exist as Python ``Gene`` objects --
calls it directly (``tools/run_config.py``: the BASELINE configs[2..4] runs).

Genomes are column indices here.  Everything random goes through the ``random`` module exactly where the reference uses
it (``random.sample`` for the K-selection sample before ``random.seed(seed)``, ``random.shuffle`` for the chunks; both
only depend on the LENGTH of the list they are given, so shuffling columns in the caller's genome order draws what the
reference draws from its organism objects).  Instances are exported and run in batches on the GPU; under
``torch.distributed`` the K candidates and the chunk samples of every draw are dealt out to the ranks (no communication
during EM; the chunks' label vectors are all-gathered and every rank applies all votes in sample order).
"""
from __future__ import annotations

import logging
import random
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import engine
from .export_device import ChunkVotes, DevicePangenome

CHUNK_BATCH = 16          # chunk instances per exporter / EM call and rank


@dataclass
class PartitionOutcome:
    labels: np.ndarray                 # per family: "P", "S<k>" / "S_", "C" (un-chunked) or "P" / "S" / "C" / "U" (chunked votes)
    K: int
    computed_K: bool
    samples: List[List[int]] = field(default_factory=list)   # chunk samples that were run (ascending columns)
    log_likelihood: Optional[float] = None
    run: Optional[object] = None       # un-chunked: the engine.NemRun (parameters for partition.run_partitioning's callers)
    timings: dict = field(default_factory=dict)


def _partition_module():
    from . import partition
    return partition


def _bcast_from_rank0(obj):
    """``obj`` of rank 0 on every rank of the process group (random draws must agree)."""
    rank, world, dist = _partition_module()._rank_world()
    if world == 1:
        return obj
    payload = [obj if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    return payload[0]


def k_candidates(dp: DevicePangenome, sample_cols: List[int], ks: List[int], free_dispersion: bool, sm_degree: float = 10):
    """([(K, M, entropy) | ({}, None, None)], nb_fam) for the K-selection sample -- run_partitioning(..., beta=0, itermax=10,
    just_log_likelihood=True) for every K (partition.py:482-514), one batched call per rank."""
    part = _partition_module()
    logger = logging.getLogger("PPanGGOLiN")
    inst = dp.export([sample_cols], sm_degree, want_levels=False)[0]
    rank, world, dist = part._rank_world()
    mine = ks[rank::world]
    runs = engine.nem_run_batch([(inst.inp, k, 0.0, free_dispersion, 10) for k in mine], flags=part._epilogue_flags(len(sample_cols)),
                                want_params=False)
    results = []
    for k, run in zip(mine, runs):
        if not run.ok:
            logger.warning("A NEM run failed while estimating the optimal number of partitions "
                           "(testing a candidate K), and this candidate was skipped. Final partitioning can still succeed.")
            results.append(({}, None, None))
            continue
        _, entropy = engine.labels_device(run.T)
        results.append((k, part._g(part._f32(run.crit[3]), "%g"), entropy))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        results = [r for p in gathered for r in p]
    return results, inst.N


def evaluate_nb_partitions_arrays(dp: DevicePangenome, cols_in_order: List[int], free_dispersion: bool = False,
                                  chunk_size: int = 500, krange=None, icl_margin: float = 0.05, sm_degree: float = 10):
    """partition.py:439-545 on arrays: (chosen K, BICs, ICLs, log-likelihoods)."""
    part = _partition_module()
    kmm = [3, 20] if krange is None else krange
    if len(cols_in_order) > chunk_size:
        sample = _bcast_from_rank0(random.sample(list(cols_in_order), chunk_size))     # partition.py:474-477, not seeded
    else:
        sample = list(cols_in_order)
    ks = list(range(kmm[0] - 1, kmm[1] + 1))
    if ks and ks[-1] > engine.MAX_K:
        logging.getLogger("PPanGGOLiN").warning(f"K candidates above {engine.MAX_K} are not supported by the B200 path "
                                                f"(NEMB200_MAX_K) and are skipped.")
        ks = [k for k in ks if k <= engine.MAX_K]
    results, nb_fam = k_candidates(dp, sorted(sample), ks, free_dispersion, sm_degree)
    return part.select_k(results, len(sample), nb_fam, free_dispersion, icl_margin)


def draw_chunks(cols_in_order: List[int], chunk_size: int, nb_sample: dict, condition: float, samples: list):
    """partition.py:810-818: shuffle the genomes and cut chunks until every genome was drawn ``condition`` times.  The
    strict '>' drops the tail of every shuffle.  Rank 0 draws, the others take its samples."""
    rank, world, _ = _partition_module()._rank_world()
    new = []
    if rank == 0:
        counts = dict(nb_sample)
        while not all(v >= condition for v in counts.values()):
            shuffled = list(cols_in_order)
            random.shuffle(shuffled)
            while len(shuffled) > chunk_size:
                new.append(shuffled[:chunk_size])
                for o in new[-1]:
                    counts[o] += 1
                shuffled = shuffled[chunk_size:]
    new = _bcast_from_rank0(new)
    for smp in new:
        samples.append(sorted(smp))
        for o in smp:
            nb_sample[o] += 1


def labels_unchunked(run, K: int) -> np.ndarray:
    """partition.py:206-231: "P", "S1".., "C", or "S_" for a tie / a maximum below 0.5."""
    lab, _ = engine.labels_device(run.T)
    names = np.array(["P"] + [f"S{i}" for i in range(1, K - 1)] + ["C", "S_"], dtype=object)   # label -1 -> "S_"
    return names[lab.cpu().numpy().astype(np.int64)]


def deal_round(ids: List[int], F: int, device, run_local, apply_codes):
    """The chunk samples ``ids`` of one draw, dealt out round-robin to the ranks in groups of CHUNK_BATCH per rank:
    ``run_local(my ids)`` -> int8 [len, F] label vectors (-1 = family absent / run failed); the vectors of a group are
    all-gathered and handed to ``apply_codes`` in SAMPLE order on every rank -- the same votes, in the same order, as a
    single process (the "undefined" rule of validate_family depends on the order)."""
    import torch
    rank, world, dist = _partition_module()._rank_world()
    for g0 in range(0, len(ids), CHUNK_BATCH * world):
        group = ids[g0:g0 + CHUNK_BATCH * world]
        codes = run_local(group[rank::world])
        if world > 1:
            per_rank = (len(group) + world - 1) // world            # pad: ranks may hold one sample less
            buf = torch.full((per_rank, F), -1, dtype=torch.int8, device=device)
            buf[:codes.shape[0]] = codes
            gathered = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(gathered, buf)
            codes = torch.stack([gathered[pos % world][pos // world] for pos in range(len(group))])   # sample order
        apply_codes(codes)


def run_chunk_round(dp: DevicePangenome, samples: List[List[int]], ids: List[int], votes: ChunkVotes, kval: int, beta: float,
                    sm_degree: float, free_dispersion: bool, chunk_size: int, stats: Optional[dict] = None):
    """One draw of chunk samples: exported and run as batches on this rank's GPU (one exporter call and one EM call per
    CHUNK_BATCH instances), votes applied on the device."""
    import torch
    part = _partition_module()
    logger = logging.getLogger("PPanGGOLiN")

    def run_local(mine):
        codes = torch.full((len(mine), dp.F), -1, dtype=torch.int8, device=dp.device)
        if not mine:
            return codes
        t0 = time.perf_counter()
        insts = dp.export([samples[i] for i in mine], sm_degree)
        t1 = time.perf_counter()
        jobs = [(it.inp, kval, it.beta_eff(beta), free_dispersion, 100) for it in insts]
        runs = engine.nem_run_batch(jobs, flags=part._epilogue_flags(chunk_size), want_params=False)
        t2 = time.perf_counter()
        for j, (it, run) in enumerate(zip(insts, runs)):
            if not run.ok:
                logger.warning("A NEM run failed for this dataset/chunk and was skipped. In chunked mode, other chunks "
                               "can still complete and final partitioning may still succeed.")
                continue
            votes.codes(run.T, kval, it.fam_rows, codes[j])
        if stats is not None:
            stats["export_s"] = stats.get("export_s", 0.0) + (t1 - t0)
            stats["em_s"] = stats.get("em_s", 0.0) + (t2 - t1)
            stats["iterations"] = stats.get("iterations", 0) + sum(r.n_iter for r in runs)
            stats["instances"] = stats.get("instances", 0) + len(runs)
        return codes

    deal_round(ids, dp.F, dp.device, run_local, votes.apply)


def partition_arrays(dp: DevicePangenome, beta: float = 2.5, sm_degree: float = 10, free_dispersion: bool = False,
                     chunk_size: int = 500, kval: int = -1, krange=None, icl_margin: float = 0.05, seed: int = 42,
                     cols_in_order: Optional[List[int]] = None, on_k=None, want_params: bool = True) -> PartitionOutcome:
    """partition.py:653-899 on arrays.  ``cols_in_order``: the genome columns in the iteration order of the caller's genome
    collection (what ``list(organisms)`` is in the reference; default 0..G-1).  ``on_k(bics, icls, lls, k)`` is called after
    K selection (the ICL plot of partition.py:547-633)."""
    logger = logging.getLogger("PPanGGOLiN")
    part = _partition_module()
    orgs = list(range(dp.G)) if cols_in_order is None else [int(c) for c in cols_in_order]
    n_org = len(orgs)
    timings, computed = {}, False
    if kval < 2:
        computed = True
        t0 = time.perf_counter()
        kval, bics, icls, lls = evaluate_nb_partitions_arrays(dp, orgs, free_dispersion, chunk_size, krange, icl_margin, sm_degree)
        timings["k_selection_s"] = time.perf_counter() - t0
        if on_k is not None and len(bics) > 0:
            on_k(bics, icls, lls, kval)
        logger.info(f"The number of partitions has been evaluated at {kval}")
    random.seed(seed)                                                   # partition.py:771: after K selection
    start = time.time()
    if chunk_size < n_org:
        votes = ChunkVotes(dp.F, n_org, chunk_size, dp.device)
        samples: List[List[int]] = []
        nb_sample = {o: 0 for o in orgs}
        condition = n_org / chunk_size
        stats = {}
        while votes.count_validated() < dp.F:
            prev = len(samples)
            draw_chunks(orgs, chunk_size, nb_sample, condition, samples)
            logger.info("Launching NEM")
            run_chunk_round(dp, samples, list(range(prev, len(samples))), votes, kval, beta, sm_degree, free_dispersion, chunk_size,
                            stats=stats)
            condition += 1
            logger.debug(f"There are {votes.count_validated()} validated families out of {dp.F} families.")
        timings.update(stats)
        timings["chunked_s"] = time.time() - start
        logger.info(f"Did {len(samples)} partitioning with chunks of size {chunk_size} among "
                    f"{n_org} genomes in {round(time.time() - start, 2)} seconds.")
        return PartitionOutcome(votes.letters(), kval, computed, samples, None, None, timings)
    # one instance over all the genomes (partition.py:866-878)
    t0 = time.perf_counter()
    cols = sorted(orgs)
    identity = cols == list(range(dp.G))
    inst = dp.export([None if identity else cols], sm_degree)[0]
    timings["export_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    run = part._run_instance(inst.inp, kval, inst.beta_eff(beta), free_dispersion, 100, part._epilogue_flags(n_org), want_params)
    timings["em_s"] = time.perf_counter() - t0
    timings["iterations"] = run.n_iter
    logger.info(f"Partitioned {n_org} genomes in {round(time.time() - start, 2)} seconds.")
    labels = np.array([""] * dp.F, dtype=object)
    if run.ok:
        labels[inst.fam_rows.cpu().numpy()] = labels_unchunked(run, kval)
    return PartitionOutcome(labels, kval, computed, [], part._g(part._f32(run.crit[3]), "%g") if run.ok else None, run, timings)
