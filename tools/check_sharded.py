#!/usr/bin/env python3
"""Multi-GPU correctness check (run under torchrun on N GPUs): the row-sharded run -- through the NVLink peer-memory
exchange and through NCCL -- must give bit-identical posteriors to the single-GPU run of the same instance, also when
the shards are uneven / start at odd offsets.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist

from ppanggolin_b200.nem import engine
from ppanggolin_b200.nem.sharded import CudaBackend, ShardedEM, shard_rows
from ppanggolin_b200.synthetic import synth_pangenome

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (N, D, K, fd) in [(3001, 300, 3, False), (2500, 130, 5, True), (4099, 1000, 3, False)]:
    sp = synth_pangenome(N, D, seed=N)
    beta = float(np.float32(sp.beta_eff(2.5)))
    full = engine.DeviceInput.from_numpy(sp.bits, D, sp.rowptr, sp.col, sp.w, device=dev)
    ref = engine.nem_run_device(full, K, beta, fd, itermax=6)
    lo, hi = shard_rows(N, world, rank)
    graph = engine.DeviceInput.from_numpy(np.zeros((1, 4), np.uint32), D, sp.rowptr, sp.col, sp.w, device=dev)
    for mode in ("peer", "nccl"):
        be = CudaBackend(full.bits[lo:hi].contiguous(), D, N, K, fd, engine.EPI_REF, graph,
                         peer_group=dist.group.WORLD if mode == "peer" else None, row_offset=lo)
        out = ShardedEM(be, N, K, rank, world, dist.group.WORLD).run(beta, itermax=6)
        same = bool(torch.equal(out.T, ref.T)) and out.n_iter == ref.n_iter
        relM = abs(out.crit[3] - ref.crit[3]) / abs(ref.crit[3])
        print(f"rank {rank} N={N} D={D} K={K} fd={fd} {mode}: identical posteriors {same} iters {out.n_iter}/{ref.n_iter} relM {relM:.1e}", flush=True)
        ok = ok and same and relM < 1e-12
        dist.barrier()
        be.plan.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
