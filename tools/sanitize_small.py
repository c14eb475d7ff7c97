"""Small shapes through every kernel (for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome

for (N, D, K, fd, beta, flags) in [(600, 100, 3, False, 2.5, engine.EPI_REF), (517, 333, 5, True, 2.5, engine.EPI_REF),
                                   (700, 129, 20, False, 0.0, engine.EPI_REF), (300, 1100, 3, False, 2.5, engine.EPI_LOGSPACE),
                                   (450, 64, 2, False, 2.5, engine.EPI_REF | engine.FULL_RECOUNT), (64, 31, 9, True, 0.0, engine.EPI_REF)]:
    sp = synth_pangenome(N, D, seed=N)
    be = sp.beta_eff(beta) if beta else 0.0
    inp = engine.DeviceInput.from_numpy(sp.bits, D, sp.rowptr, sp.col, sp.w)
    r = engine.nem_run_device(inp, K, be, fd, 12, flags=flags)
    lab, ent = engine.labels_device(r.T)
    h = engine.nem_run_host(sp.bits, D, sp.rowptr, sp.col, sp.w, K, be, fd, 12, flags=flags)
    print(N, D, K, fd, "status", r.status, "iters", r.n_iter, "host==device", bool(np.array_equal(h.T, r.T.cpu().numpy())), flush=True)
sp = synth_pangenome(500, 90, seed=1)
inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
runs = engine.nem_run_batch([(inp, k, 0.0, False, 10) for k in (2, 3, 7, 20)] + [(inp, 3, sp.beta_eff(2.5), True, 20)])
print("batch", [r.n_iter for r in runs], flush=True)
torch.cuda.synchronize()
print("done")
