#!/usr/bin/env python3
"""Wall time of the public API partition() on in-memory pangenome objects: ours (duck-typed objects, GPU) or, with
--reference (only where /root/reference exists), the reference's own partition() on its own objects, fork pool over all
host cores.  Same generator (tests/golden/make_golden.py::make_layouts), same arguments.
usage: tools/bench_partition_api.py [--reference] [--fam 5000] [--org 400] [--chunk 100]"""
import argparse
import json
import os
import random
import sys
import tempfile
import time
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tests" / "golden"))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--reference", action="store_true")
ap.add_argument("--fam", type=int, default=5000)
ap.add_argument("--org", type=int, default=400)
ap.add_argument("--chunk", type=int, default=100)
a = ap.parse_args()

import make_golden as mg
layouts = mg.make_layouts(a.fam, a.org, 31)
cases = [("fixedK", dict(kval=3, chunk_size=10**6)), ("autoK", dict(kval=-1, chunk_size=10**6, krange=[3, 20])),
         ("chunked", dict(kval=3, chunk_size=a.chunk))]
out = {"families": a.fam, "genomes": a.org, "impl": "reference" if a.reference else "b200"}
if a.reference:
    ref_partition = mg.import_reference_partition()
    t = time.time(); pan = mg.build_reference_pangenome(layouts, a.fam); out["build_objects_s"] = time.time() - t
    cpu = os.cpu_count()
    out["cores"] = cpu
    for name, kw in cases:
        for f in pan.gene_families:
            f.partition = ""
        pan.status["partitioned"] = "No"
        ref_partition.samples = []
        random.seed(1234)
        with tempfile.TemporaryDirectory() as td:
            t = time.time()
            ref_partition.partition(pan, output=Path(td), tmpdir=Path(td), cpu=cpu, disable_bar=True, seed=42, beta=2.5, **kw)
            out[name + "_s"] = time.time() - t
        out[name + "_counts"] = dict(Counter(f.partition[:1] for f in pan.gene_families))
        out[name + "_instances"] = len(ref_partition.samples)
else:
    import minipan
    import torch
    from ppanggolin_b200.nem import partition as part
    t = time.time(); pan = minipan.build_from_layouts(layouts, a.fam); out["build_objects_s"] = time.time() - t
    torch.zeros(1).cuda()
    for rep in range(2):            # first round warms CUDA / the allocator cache
        for name, kw in cases:
            for f in pan.gene_families:
                f.partition = ""
            pan.status["partitioned"] = "No"
            part.samples = []
            random.seed(1234)
            with tempfile.TemporaryDirectory() as td:
                torch.cuda.synchronize(); t = time.time()
                part.partition(pan, output=Path(td), tmpdir=Path(td), cpu=1, disable_bar=True, seed=42, beta=2.5, **kw)
                torch.cuda.synchronize(); out[name + "_s"] = time.time() - t
            out[name + "_counts"] = dict(Counter(f.partition[:1] for f in pan.gene_families))
            out[name + "_instances"] = len(part.samples)
            out[name + "_K"] = pan.parameters["partition"]["# final nb of partitions"]
print(json.dumps(out))
