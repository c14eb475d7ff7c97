#!/usr/bin/env python3
"""partition() in chunked mode and with automatic K under a process group: the chunk samples / K candidates are dealt out
to the ranks (one GPU each); every rank must end with the same partition (random genome samples are drawn on rank 0 and broadcast; a single process
draws other samples, so its class counts are close, not identical).
  python tools/check_chunked_ranks.py                      # single process: prints the reference digest
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_chunked_ranks.py"""
import hashlib
import json
import os
import random
import sys
import tempfile
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tests" / "golden"))
import torch
import torch.distributed as dist

import make_golden as mg
import minipan
from ppanggolin_b200.nem import partition as part

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
junk = [object() for _ in range(rank * 100003)]   # different heap layouts per rank: set iteration orders of objects differ
layouts = mg.make_layouts(3000, 240, 31)
pan = minipan.build_from_layouts(layouts, 3000)
part.SHARD_MIN_BYTES = 1 << 62
cases = [("chunked", dict(kval=3, chunk_size=60)), ("autoK", dict(kval=-1, chunk_size=10**6, krange=[3, 12])),
         ("fixedK on one GPU per rank", dict(kval=3, chunk_size=10**6)), ("fixedK row-sharded", dict(kval=3, chunk_size=10**6))]
for name, kw in cases:
    part.SHARD_MIN_BYTES = 0 if name.endswith("row-sharded") else 1 << 62
    for f in pan.gene_families:
        f.partition = ""
    pan.status["partitioned"] = "No"
    part.samples = []
    random.seed(1234)
    with tempfile.TemporaryDirectory() as td:
        part.partition(pan, output=Path(td), tmpdir=Path(td), cpu=1, disable_bar=True, seed=42, beta=2.5, **kw)
    labels = "".join(f"{f.name}:{f.partition};" for f in sorted(pan.gene_families, key=lambda f: f.name))
    digest = hashlib.sha256(labels.encode()).hexdigest()[:16]
    print("CHECK " + json.dumps({"world": world, "rank": rank, "case": name, "digest": digest}), flush=True)
    print(f"world {world} rank {rank} {name}: instances {len(part.samples)} K {pan.parameters['partition']['# final nb of partitions']} "
          f"counts {dict(Counter(f.partition[:1] for f in pan.gene_families))} digest {digest}", flush=True)
from ppanggolin_b200.nem import rarefaction as raref
for f in pan.gene_families:
    f.partition = ""
random.seed(99)
with tempfile.TemporaryDirectory() as td:
    recs = raref.make_rarefaction_curve(pan, output=None, tmpdir=Path(td), depth=2, min_sampling=20, max_sampling=40, chunk_size=30,
                                        kval=3, disable_bar=True)
text = ";".join(",".join(f"{k}={r[k]}" for k in sorted(r)) for r in recs)
print("CHECK " + json.dumps({"world": world, "rank": rank, "case": "rarefaction", "digest": hashlib.sha256(text.encode()).hexdigest()[:16]}), flush=True)
print(f"world {world} rank {rank} rarefaction: {len(recs)} samples digest {hashlib.sha256(text.encode()).hexdigest()[:16]}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
