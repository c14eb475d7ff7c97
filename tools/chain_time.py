import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from ppanggolin_b200.nem import engine
rng = np.random.default_rng(0)
for n in (60_000, 300_000, 1_000_000):
    d = (-np.abs(rng.normal(120, 60, n))).astype(np.float32)
    for name, g in (("g zero", np.zeros(n, np.float32)), ("g small", (rng.random(n) * 1e-2).astype(np.float32))):
        for par in (False, True):
            engine.float_chain(d, g, par)
            t = time.perf_counter(); r = engine.float_chain(d, g, par); dt = time.perf_counter() - t
            print(n, name, "parallel" if par else "sequential", f"{dt*1e3:.2f} ms (incl. H2D)", r)
