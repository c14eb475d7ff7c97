import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ppanggolin_b200.nem import engine
from ppanggolin_b200.synthetic import synth_pangenome
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sp = synth_pangenome(50000, 500, seed=43)
inp = engine.DeviceInput.from_numpy(sp.bits, sp.D, sp.rowptr, sp.col, sp.w)
engine.nem_run_device(inp, K, 0.0, False, 10)
torch.cuda.synchronize(); t = time.perf_counter()
r = engine.nem_run_device(inp, K, 0.0, False, 10)
torch.cuda.synchronize(); print("K", K, "iters", r.n_iter, "wall", time.perf_counter() - t, "em_ms", r.elapsed_ms)
