"""Multi-GPU parity gates (skipped on boxes with fewer than 2 GPUs): one process per GPU under torchrun.

* row-sharded EM (nem/sharded.py) through the NVLink peer-memory exchange and through NCCL: posteriors bit-identical to
  the single-GPU run, uneven shards at odd offsets (tools/check_sharded.py);
* the instance-parallel modes of partition() -- chunk samples, K candidates, rarefaction samples dealt over the ranks --
  end with the same partition on every rank, and the row-sharded fixed-K partition equals the one-GPU-per-rank one
  (tools/check_chunked_ranks.py)."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script, n, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tools" / script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ))


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_row_sharded_em_is_bit_identical_on_two_gpus():
    r = _torchrun("check_sharded.py", min(_gpus(), 3))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if "identical posteriors" in l]
    assert lines and all("identical posteriors True" in l for l in lines)


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
def test_instance_parallel_partition_agrees_across_ranks():
    r = _torchrun("check_chunked_ranks.py", 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    recs = [json.loads(l[6:]) for l in r.stdout.splitlines() if l.startswith("CHECK ")]
    by_case = {}
    for x in recs:
        by_case.setdefault(x["case"], {})[x["rank"]] = x["digest"]
    assert len(by_case) == 5, by_case
    for case, d in by_case.items():
        assert len(d) == 2 and len(set(d.values())) == 1, (case, d)
    assert by_case["fixedK on one GPU per rank"][0] == by_case["fixedK row-sharded"][0]
