"""The C-ABI library loads and exports every symbol include/nemb200.h declares; host-only entry points work; compute
entry points fail loudly (never fall back) without a GPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from ppanggolin_b200.nem import engine

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "nemb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nemb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 15
    raw = ctypes.CDLL(str(engine.LIB_PATH))
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/nemb200.h but not exported by libnemb200.so"
        assert s in engine.SIGNATURES, f"{s} has no ctypes signature in engine.SIGNATURES"
    assert set(engine.SIGNATURES) == set(syms)
    assert engine.lib().nemb200_abi_version() == 2


def test_gs_schedule_levels_respect_dependencies():
    rng = np.random.default_rng(0)
    N = 400
    rows = [sorted(set(rng.integers(0, N, rng.integers(0, 5)).tolist())) for _ in range(N)]
    rowptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int32)
    col = np.array([j for r in rows for j in r], np.int32)
    lp, order, nl = engine.gs_schedule(rowptr, col)
    level = np.empty(N, int)
    for l in range(nl):
        level[order[lp[l]:lp[l + 1]]] = l
    assert sorted(order.tolist()) == list(range(N)) and lp[-1] == N
    for i in range(N):
        for j in rows[i]:
            if j < i:
                assert level[j] < level[i]
        want = max([level[j] + 1 for j in rows[i] if j < i], default=0)
        assert level[i] == want
    # a pure chain in index order is the worst case: one level per row
    rowptr = np.arange(0, N + 1, dtype=np.int32); rowptr[0] = 0
    col = np.maximum(np.arange(N) - 1, 0).astype(np.int32)
    assert engine.gs_schedule(rowptr, col)[2] == N


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    bits = np.zeros((4, 4), np.uint32)
    with pytest.raises(engine.NemB200Error):
        engine.nem_run_host(bits, 100, None, None, None, 3, 0.0)
    res = engine._Result()
    T = np.zeros((4, 3), np.float32)
    rc = engine.lib().nemb200_nem_host(bits.ctypes.data, 4, 100, 4, 0, 0, 0, 3, 0.0, 0, 10, 0.01, 0, T.ctypes.data, 0, 0, 0, ctypes.byref(res))
    assert rc == -2 and b"cuda" in engine.lib().nemb200_last_error().lower()


def test_bad_arguments_are_rejected():
    h = ctypes.c_void_p()
    assert engine.lib().nemb200_plan_create(ctypes.byref(h), 10, 10, 100, 3, 3, 0, 0) == -1      # stride not multiple of 4
    assert engine.lib().nemb200_plan_create(ctypes.byref(h), 10, 10, 100, 4, 33, 0, 0) == -1     # K > 32
    assert engine.lib().nemb200_plan_create(ctypes.byref(h), 10, 10, 200, 4, 3, 0, 0) == -1      # stride too short for D


def test_reference_nem_symbol_is_exported_with_reference_error_codes(tmp_path):
    """include/nem_stats_b200.h: `int nem(...)` with the signature of NEM/nem_exe.h:23-38 (what nem_stats.pyx binds)."""
    import ctypes as C
    L = C.CDLL(str(engine.LIB_PATH))
    assert hasattr(L, "nem")
    L.nem.restype = C.c_int
    L.nem.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_float, C.c_char_p, C.c_float, C.c_char_p, C.c_int, C.c_int,
                      C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    base = str(tmp_path / "nem_file").encode()
    args = lambda **kw: [kw.get("f", base), kw.get("k", 3), kw.get("algo", b"nem"), 1.0, b"clas", 0.01, b"fuzzy", 100, 1, b"bern",
                         b"pk", kw.get("disp", b"sk_"), kw.get("init", 2), base + b"_init_3.m", base + b"_3", 42]
    assert L.nem(*args(algo=b"gem")) == 2          # EXIT_E_ARGS: only what PPanGGOLiN passes is supported
    assert L.nem(*args(init=1)) == 2
    assert L.nem(*args()) == 3                     # EXIT_E_FILE: no input files
    assert b"can't open" in (tmp_path / "nem_file_3.stderr").read_bytes()
