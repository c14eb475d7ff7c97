import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome
N, D, K, fd, beta = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4])), float(sys.argv[5])
sp = synth_pangenome(N, D, seed=43)
be = sp.beta_eff(beta) if beta else 0.0
inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
engine.nem_run_device(inp, K, be, fd, 100)
torch.cuda.synchronize(); t = time.perf_counter()
r = engine.nem_run_device(inp, K, be, fd, 100)
torch.cuda.synchronize(); print("N", N, "D", D, "K", K, "fd", fd, "iters", r.n_iter, "levels", r.n_levels, "wall", time.perf_counter() - t, "em_ms", r.elapsed_ms)
