"""Row-sharded EM driver under gloo, world_size 2, on CPU: the sharded run must reproduce the single-process run bit
for bit (the M-step statistics are all-reduced, the log-densities all-gathered, the sweep replicated)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(rank, world, port, N, D, K, beta, iters, out):
    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from npbackend import NumpyBackend
    from ppanggolin_b200.nem.sharded import ShardedEM, shard_rows
    from ppanggolin_b200.synthetic import synth_pangenome
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    group = None
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        group = dist.group.WORLD
    sp = synth_pangenome(N, D, seed=3)
    lo, hi = shard_rows(N, world, rank)
    be = NumpyBackend(sp.X[lo:hi], D, N, K, False, (sp.rowptr, sp.col, sp.w))
    em = ShardedEM(be, N, K, rank, world, group)
    b = float(np.float32(sp.beta_eff(beta)))
    em.start(b)
    flags = [em.iterate(b) for _ in range(iters)]
    np.save(f"{out}/T_{world}_{rank}.npy", be.T[em.cur].numpy())
    np.save(f"{out}/mu_{world}_{rank}.npy", be.mu)
    np.save(f"{out}/flags_{world}_{rank}.npy", np.array(flags))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [301, 400])   # uneven and even shards (all_gather list path / all_gather_into_tensor path)
def test_two_rank_gloo_run_equals_single_process(tmp_path, N):
    D, K, beta, iters = 40, 3, 2.5, 3
    _run(0, 1, _free_port(), N, D, K, beta, iters, str(tmp_path))
    port = _free_port()
    mp.start_processes(_run, args=(2, port, N, D, K, beta, iters, str(tmp_path)), nprocs=2, join=True, start_method="fork")
    T1 = np.load(tmp_path / "T_1_0.npy")
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"T_2_{r}.npy"), T1)              # replicated, identical posteriors
        assert np.array_equal(np.load(tmp_path / f"mu_2_{r}.npy"), np.load(tmp_path / "mu_1_0.npy"))
        assert np.allclose(np.load(tmp_path / f"flags_2_{r}.npy"), np.load(tmp_path / "flags_1_0.npy"))
    assert (T1.argmax(1) == 0).sum() > 0 and (T1.argmax(1) == K - 1).sum() > 0


def _chunk_votes(rank, world, port, F, n_samples, out):
    """Chunk samples dealt out to the ranks (workflow.deal_round) with a deterministic stand-in for the GPU runs."""
    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import npref
    from ppanggolin_b200.nem import workflow
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    n_org, chunk_size = 40, 10

    def fake_batch(ids):   # label vector of chunk i: depends only on i, some families absent, some chunks failed (-1 everywhere)
        res = torch.full((len(ids), F), -1, dtype=torch.int8)
        for j, i in enumerate(ids):
            g = torch.Generator().manual_seed(1000 + i)
            present = torch.rand(F, generator=g) < 0.7
            code = torch.randint(0, 3, (F,), generator=g).to(torch.int8)
            if i % 11 != 5:
                res[j] = torch.where(present, code, torch.tensor(-1, dtype=torch.int8))
        return res

    votes = np.zeros((F, 4), np.int64)
    validated = np.zeros(F, bool)
    old = workflow.CHUNK_BATCH
    workflow.CHUNK_BATCH = 3          # several groups, the last one ragged
    try:
        workflow.deal_round(list(range(n_samples)), F, torch.device("cpu"), fake_batch,
                            lambda codes: npref.apply_votes_ref(codes.numpy(), votes, validated, n_org, chunk_size))
    finally:
        workflow.CHUNK_BATCH = old
    np.save(f"{out}/votes_{world}_{rank}.npy", votes)
    np.save(f"{out}/validated_{world}_{rank}.npy", validated)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [17, 24])
def test_chunk_samples_dealt_to_two_ranks_give_the_same_votes(tmp_path, n_samples):
    """Independent chunk instances on two ranks (gloo): every rank ends with the votes and the validation state of the
    single-process run, applied in the same sample order (the "undefined" rule of validate_family depends on it)."""
    F = 300
    _chunk_votes(0, 1, _free_port(), F, n_samples, str(tmp_path))
    port = _free_port()
    mp.spawn(_chunk_votes, args=(2, port, F, n_samples, str(tmp_path)), nprocs=2, join=True)
    v1 = np.load(tmp_path / "votes_1_0.npy"); ok1 = np.load(tmp_path / "validated_1_0.npy")
    assert v1.sum() > 0 and ok1.any()
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"votes_2_{r}.npy"), v1)
        assert np.array_equal(np.load(tmp_path / f"validated_2_{r}.npy"), ok1)


class _Org:
    def __init__(self, name):
        self.name = name


def _sample_sync(rank, world, port, out):
    """Random genome samples under a process group: rank 0's draw reaches every rank, by genome name."""
    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import json
    import random
    from collections import Counter
    from ppanggolin_b200.nem import partition as ppp
    from ppanggolin_b200.nem import workflow
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orgs = [_Org(f"g{i:03d}") for i in range(57)]
    random.seed(100 + rank)                   # the ranks' own random streams and collection orders differ on purpose
    random.shuffle(orgs)
    chosen = ppp._same_on_all_ranks(set(random.sample(orgs, 20)), orgs)
    assert all(o in orgs for o in chosen)     # this rank's own objects, not unpickled copies
    # the array workflow draws genome COLUMNS: the K-selection sample and the chunk samples of rank 0 reach every rank
    cols = list(range(57))
    random.shuffle(cols)
    k_sample = workflow._bcast_from_rank0(random.sample(cols, 20))
    samples = []
    counts = {c: 0 for c in cols}
    workflow.draw_chunks(cols, 10, counts, 2, samples)
    rec = {"chosen": sorted(o.name for o in chosen), "k_sample": sorted(k_sample), "samples": samples,
           "counts": sorted(counts.items())}
    with open(f"{out}/sync_{rank}.json", "w") as f:
        json.dump(rec, f)
    with ppp.local_instances():               # nested callers switch the dealing-out off
        assert ppp._rank_world()[:2] == (0, 1)
    assert ppp._rank_world()[:2] == (rank, world)
    dist.barrier()
    dist.destroy_process_group()


def test_random_genome_samples_are_the_same_on_every_rank(tmp_path):
    import json
    mp.spawn(_sample_sync, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a = json.load(open(tmp_path / "sync_0.json")); b = json.load(open(tmp_path / "sync_1.json"))
    assert a == b
    assert len(a["chosen"]) == 20 and len(a["k_sample"]) == 20
    assert len(a["samples"]) >= 10 and all(len(s) == 10 and s == sorted(s) for s in a["samples"])
    assert all(c >= 2 for _, c in a["counts"])
