"""Python side of the C ABI in ``include/nemb200.h`` (``libnemb200.so``: hand-written sm_100a CUDA).

PyTorch is plumbing here: device memory (tensors), the current CUDA stream and ``torch.distributed``.  The compute
entry points are registered as ``torch.library`` custom ops (``ppanggolin_b200::nem_run``) whose implementation passes
raw device pointers to the C ABI through ctypes.  There is no CPU fallback: if the shared library is missing or no
CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

import numpy as np

PKG_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = PKG_DIR / "libnemb200.so"

EPI_REF = 0
EPI_LOGSPACE = 1
JACOBI = 2
FULL_RECOUNT = 4
PEER_EXCHANGE = 8
TIME_DENSITY = 16
MAX_K = 32
MAX_PEERS = 8          # kMaxPeers of the peer-memory exchange (one NVSwitch box)
PEER_OFFSETS = 16      # NEMB200_PEER_OFFSETS

STS_OK, STS_EMPTYCLASS, STS_INFINITE = 0, 1, 2


class NemB200Error(RuntimeError):
    pass


class _Result(ctypes.Structure):
    _fields_ = [("crit", ctypes.c_double * 5), ("n_iter", ctypes.c_int32), ("converged", ctypes.c_int32),
                ("status", ctypes.c_int32), ("n_levels", ctypes.c_int32), ("elapsed_ms", ctypes.c_float),
                ("em_ms", ctypes.c_float), ("density_ms", ctypes.c_float), ("density_launches", ctypes.c_int32)]


_lib = None
_vp, _i64, _i32, _u32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32, ctypes.c_float

# every symbol include/nemb200.h declares, with its ctypes signature
SIGNATURES = {
    "nemb200_abi_version": (ctypes.c_int, []),
    "nemb200_last_error": (ctypes.c_char_p, []),
    "nemb200_device_count": (ctypes.c_int, []),
    "nemb200_release_cached_memory": (ctypes.c_int, []),
    "nemb200_nem_host": (ctypes.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _f32, _i32, _i32, _f32, _u32,
                                        _vp, _vp, _vp, _vp, ctypes.POINTER(_Result)]),
    "nemb200_nem_device": (ctypes.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _i32,
                                          _f32, _u32, _vp, _vp, _vp, _vp, ctypes.POINTER(_Result), _vp]),
    "nemb200_nem_device_batch": (ctypes.c_int, [_i32] + [_vp] * 14 + [_f32, _u32] + [_vp] * 5),
    "nemb200_gs_schedule": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "nemb200_plan_create": (ctypes.c_int, [ctypes.POINTER(_vp), _i64, _i64, _i64, _i64, _i32, _i32, _u32]),
    "nemb200_plan_destroy": (None, [_vp]),
    "nemb200_plan_timing": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_f32), ctypes.POINTER(_i32)]),
    "nemb200_init_params": (ctypes.c_int, [_vp, _vp]),
    "nemb200_mstep_accumulate": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "nemb200_stats_ptrs": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                          ctypes.POINTER(_vp), ctypes.POINTER(_i64)]),
    "nemb200_stats_blocks": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_i64), ctypes.POINTER(_vp), ctypes.POINTER(_i64)]),
    "nemb200_mstep_finalize": (ctypes.c_int, [_vp, _vp]),
    "nemb200_density": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "nemb200_peer_export": (ctypes.c_int, [_vp, _vp, _vp]),
    "nemb200_peer_attach": (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "nemb200_peer_logf": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "nemb200_density_peers": (ctypes.c_int, [_vp, _vp, _i64, _vp]),
    "nemb200_peer_wait_densities": (ctypes.c_int, [_vp, _vp]),
    "nemb200_nem_sharded": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _i32, _f32, _vp,
                                           ctypes.POINTER(_Result), _vp]),
    "nemb200_sweep": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _vp]),
    "nemb200_plan_graph_changed": (ctypes.c_int, [_vp]),
    "nemb200_selftest_float_chain": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "nemb200_selftest_double_chain": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "nemb200_export_samples": (ctypes.c_int, [_vp, _i32, _vp, _f32, _vp, _i32, _i32, _vp]),
    "nemb200_chunk_codes": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "nemb200_apply_votes": (ctypes.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, ctypes.c_double, _i32, _vp]),
    "nemb200_subset_counts": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "nemb200_read_flags": (ctypes.c_int, [_vp, ctypes.POINTER(_f32), ctypes.POINTER(_i32), _vp]),
    "nemb200_criteria": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _f32, ctypes.POINTER(ctypes.c_double * 5), _vp]),
    "nemb200_get_params": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.POINTER(_i32), _vp]),
    "nemb200_labels": (ctypes.c_int, [_vp, _i64, _i32, _vp, ctypes.POINTER(ctypes.c_double), _vp]),
}


def lib():
    """Load ``libnemb200.so`` (built in-tree by ``__graft_entry__.build()``).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.is_file():
            raise NemB200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the NEM path)")
        L = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().nemb200_last_error()
        raise NemB200Error(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def _torch():
    import torch
    return torch


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise NemB200Error("no CUDA device: the NEM path runs only on the GPU (no CPU fallback)")
    return torch


def _ptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())


def _stream() -> int:
    return int(_torch().cuda.current_stream().cuda_stream)


@dataclass
class NemRun:
    """Outcome of one NEM instance (what nem() leaves in .uf/.mf, nem_exe.c:1601-1786)."""
    T: object                      # (N, K) float32 posteriors: torch tensor (device run) or numpy array (host run)
    mu: np.ndarray                 # (K, D) float32 in {0, 0.5, 1}
    eps: np.ndarray                # (K, D) float32
    pk: np.ndarray                 # (K,) float32
    crit: list                     # U, D, L, M, Z
    n_iter: int
    converged: bool
    status: int
    n_levels: int
    elapsed_ms: float
    em_ms: float = 0.0
    density_ms: float = 0.0        # TIME_DENSITY: device time of the density launches of the EM iterations, and their number
    density_launches: int = 0

    @property
    def ok(self) -> bool:
        return self.status == STS_OK


def gs_schedule(rowptr: np.ndarray, col: np.ndarray):
    """Level schedule of the in-place (Gauss-Seidel) sweep; host helper."""
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    N = len(rowptr) - 1
    lp = np.zeros(N + 1, np.int32)
    rows = np.zeros(N, np.int32)
    nl = lib().nemb200_gs_schedule(N, rowptr.ctypes.data, col.ctypes.data if len(col) else 0, lp.ctypes.data, rows.ctypes.data)
    if nl < 0:
        _check(nl, "nemb200_gs_schedule")
    return lp[: nl + 1].copy(), rows, nl


def nem_run_host(bits: np.ndarray, D: int, rowptr, col, w, K: int, beta: float, free_dispersion: bool = False,
                 itermax: int = 100, cv_th: float = 0.01, flags: int = EPI_REF) -> NemRun:
    """Whole run through the host-buffer entry (host<->device copies inside): the drop-in for ``nem_stats.nem``."""
    _require_cuda()
    bits = np.ascontiguousarray(bits).view(np.uint32)
    N, Ws = bits.shape
    T = np.empty((N, K), np.float32); mu = np.empty((K, D), np.float32); eps = np.empty((K, D), np.float32)
    pk = np.empty(K, np.float32)
    res = _Result()
    if beta != 0:
        rowptr = np.ascontiguousarray(rowptr, np.int32); col = np.ascontiguousarray(col, np.int32)
        w = np.ascontiguousarray(w, np.float32)
        rp, cp, wp = rowptr.ctypes.data, col.ctypes.data, w.ctypes.data
    else:
        rp = cp = wp = 0
    rc = lib().nemb200_nem_host(bits.ctypes.data, N, D, Ws, rp, cp, wp, K, float(np.float32(beta)), int(free_dispersion),
                                itermax, cv_th, flags, T.ctypes.data, mu.ctypes.data, eps.ctypes.data, pk.ctypes.data,
                                ctypes.byref(res))
    _check(rc, "nemb200_nem_host")
    return NemRun(T, mu, eps, pk, list(res.crit), res.n_iter, bool(res.converged), res.status, res.n_levels, res.elapsed_ms,
                  res.em_ms, res.density_ms, res.density_launches)


@dataclass
class DeviceInput:
    """One NEM instance resident in HBM."""
    bits: object                   # torch int32 (N, Ws) on cuda
    D: int
    rowptr: Optional[object] = None
    col: Optional[object] = None
    w: Optional[object] = None
    sched_ptr: Optional[object] = None
    sched_rows: Optional[object] = None
    n_levels: int = 1

    @property
    def N(self) -> int:
        return int(self.bits.shape[0])

    @staticmethod
    def from_tensor(bits_t, D: int, rowptr=None, col=None, w=None) -> "DeviceInput":
        """Matrix already resident on the device (export_arrays.PangenomeArrays.instance); CSR as numpy arrays."""
        torch = _require_cuda()
        di = DeviceInput(bits_t.contiguous(), D)
        if rowptr is not None and len(col) > 0:
            dev = bits_t.device
            lp, rows, nl = gs_schedule(rowptr, col)
            di.rowptr = torch.from_numpy(np.ascontiguousarray(rowptr, np.int32)).to(dev)
            di.col = torch.from_numpy(np.ascontiguousarray(col, np.int32)).to(dev)
            di.w = torch.from_numpy(np.ascontiguousarray(w, np.float32)).to(dev)
            di.sched_ptr = torch.from_numpy(lp).to(dev)
            di.sched_rows = torch.from_numpy(rows).to(dev)
            di.n_levels = nl
        return di

    @staticmethod
    def from_numpy(bits: np.ndarray, D: int, rowptr=None, col=None, w=None, device="cuda") -> "DeviceInput":
        torch = _require_cuda()
        b = torch.from_numpy(np.ascontiguousarray(bits).view(np.int32)).to(device)
        di = DeviceInput(b, D)
        if rowptr is not None and len(col) > 0:
            lp, rows, nl = gs_schedule(rowptr, col)
            di.rowptr = torch.from_numpy(np.ascontiguousarray(rowptr, np.int32)).to(device)
            di.col = torch.from_numpy(np.ascontiguousarray(col, np.int32)).to(device)
            di.w = torch.from_numpy(np.ascontiguousarray(w, np.float32)).to(device)
            di.sched_ptr = torch.from_numpy(lp).to(device)
            di.sched_rows = torch.from_numpy(rows).to(device)
            di.n_levels = nl
        return di


def _nem_run_device_impl(bits, rowptr, col, w, sched_ptr, sched_rows, D: int, n_levels: int, K: int, beta: float,
                         free_disp: int, itermax: int, cv_th: float, flags: int, want_params: bool = True):
    torch = _torch()
    N, Ws = bits.shape
    dev = bits.device
    T = torch.empty((N, K), dtype=torch.float32, device=dev)
    kd = (K, D) if want_params else (0, 0)
    mu = torch.empty(kd, dtype=torch.float32, device=dev)
    eps = torch.empty(kd, dtype=torch.float32, device=dev)
    pk = torch.empty((K if want_params else 0,), dtype=torch.float32, device=dev)
    res = _Result()
    with torch.cuda.device(dev):
        rc = lib().nemb200_nem_device(_ptr(bits), N, D, Ws, _ptr(rowptr), _ptr(col), _ptr(w), _ptr(sched_ptr),
                                      _ptr(sched_rows), n_levels, K, beta, free_disp, itermax, cv_th, flags, _ptr(T),
                                      _ptr(mu) if want_params else 0, _ptr(eps) if want_params else 0,
                                      _ptr(pk) if want_params else 0, ctypes.byref(res), _stream())
    _check(rc, "nemb200_nem_device")
    stats = torch.tensor(list(res.crit) + [res.n_iter, res.converged, res.status, res.n_levels, res.elapsed_ms, res.em_ms,
                                           res.density_ms, res.density_launches],
                         dtype=torch.float64)
    return T, mu, eps, pk, stats


_op_registered = False


def _register_op():
    """``ppanggolin_b200::nem_run`` as a torch.library custom op (opaque to autograd/compile; CUDA only)."""
    global _op_registered
    if _op_registered:
        return
    torch = _torch()
    torch.library.define(
        "ppanggolin_b200::nem_run",
        "(Tensor bits, Tensor? rowptr, Tensor? col, Tensor? w, Tensor? sched_ptr, Tensor? sched_rows, int D, "
        "int n_levels, int K, float beta, int free_disp, int itermax, float cv_th, int flags, bool want_params=True) "
        "-> (Tensor, Tensor, Tensor, Tensor, Tensor)")

    @torch.library.impl("ppanggolin_b200::nem_run", "CUDA")
    def nem_run(bits, rowptr, col, w, sched_ptr, sched_rows, D, n_levels, K, beta, free_disp, itermax, cv_th, flags,
                want_params=True):
        return _nem_run_device_impl(bits, rowptr, col, w, sched_ptr, sched_rows, D, n_levels, K, beta, free_disp,
                                    itermax, cv_th, flags, want_params)

    _op_registered = True


def nem_run_device(inp: DeviceInput, K: int, beta: float, free_dispersion: bool = False, itermax: int = 100,
                   cv_th: float = 0.01, flags: int = EPI_REF, want_params: bool = True) -> NemRun:
    """Whole run with the instance already resident in HBM (the custom op).  ``want_params=False`` skips the K x D
    centres / dispersions (callers that only need posteriors and criteria: K selection, chunk votes)."""
    torch = _require_cuda()
    if not (1 <= K <= MAX_K):
        raise NemB200Error(f"K={K} outside 1..{MAX_K}")
    _register_op()
    beta32 = float(np.float32(beta))
    use_graph = beta32 != 0.0 and inp.rowptr is not None
    if beta32 != 0.0 and inp.rowptr is None:
        raise NemB200Error("beta != 0 needs the neighbour graph")
    T, mu, eps, pk, stats = torch.ops.ppanggolin_b200.nem_run(
        inp.bits, inp.rowptr if use_graph else None, inp.col if use_graph else None, inp.w if use_graph else None,
        inp.sched_ptr if use_graph else None, inp.sched_rows if use_graph else None, inp.D,
        inp.n_levels if use_graph else 1, K, beta32, int(free_dispersion), itermax, cv_th, flags, want_params)
    s = stats.tolist()
    if want_params:
        mu, eps, pk = mu.cpu().numpy(), eps.cpu().numpy(), pk.cpu().numpy()
    else:
        mu = eps = pk = None
    return NemRun(T, mu, eps, pk, s[:5], int(s[5]), bool(s[6]), int(s[7]), int(s[8]), float(s[9]), float(s[10]), float(s[11]), int(s[12]))


def nem_run_batch(instances, cv_th: float = 0.01, flags: int = EPI_REF, want_params: bool = True):
    """Independent NEM instances in one call (``nemb200_nem_device_batch``): the K-selection sweep or chunk samples.

    ``instances``: list of ``(DeviceInput, K, beta, free_dispersion, itermax)``.  Returns a list of :class:`NemRun`
    (``mu`` / ``eps`` / ``pk`` are None with ``want_params=False``)."""
    torch = _require_cuda()
    n = len(instances)
    if n == 0:
        return []
    dev = instances[0][0].bits.device
    outs = []
    P = ctypes.c_void_p * n
    bits, rp, cl, ww, sp, sr = P(), P(), P(), P(), P(), P()
    Tp, mup, epsp, pkp = P(), P(), P(), P()
    Ns, Ds, Ws = (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)()
    nl, Ks, fds, its = (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)()
    betas = (ctypes.c_float * n)()
    res = (_Result * n)()
    for i, (inp, K, beta, fd, itermax) in enumerate(instances):
        if not (1 <= K <= MAX_K):
            raise NemB200Error(f"K={K} outside 1..{MAX_K}")
        beta32 = float(np.float32(beta))
        g = beta32 != 0.0
        if g and inp.rowptr is None:
            raise NemB200Error("beta != 0 needs the neighbour graph")
        N, W = inp.bits.shape
        T = torch.empty((N, K), dtype=torch.float32, device=dev)
        mu = eps = pk = None
        if want_params:
            mu = torch.empty((K, inp.D), dtype=torch.float32, device=dev)
            eps = torch.empty((K, inp.D), dtype=torch.float32, device=dev)
            pk = torch.empty((K,), dtype=torch.float32, device=dev)
        outs.append((T, mu, eps, pk))
        bits[i], Ns[i], Ds[i], Ws[i] = _ptr(inp.bits), N, inp.D, W
        rp[i], cl[i], ww[i] = (_ptr(inp.rowptr), _ptr(inp.col), _ptr(inp.w)) if g else (0, 0, 0)
        sp[i], sr[i], nl[i] = (_ptr(inp.sched_ptr), _ptr(inp.sched_rows), inp.n_levels) if g else (0, 0, 1)
        Ks[i], betas[i], fds[i], its[i] = K, beta32, int(bool(fd)), itermax
        Tp[i], mup[i], epsp[i], pkp[i] = _ptr(T), _ptr(mu), _ptr(eps), _ptr(pk)
    torch.cuda.current_stream().synchronize()   # the instances run on their own streams: inputs must be complete
    addr = lambda a: ctypes.cast(a, ctypes.c_void_p)
    with torch.cuda.device(dev):
        rc = lib().nemb200_nem_device_batch(n, addr(bits), addr(Ns), addr(Ds), addr(Ws), addr(rp), addr(cl), addr(ww), addr(sp),
                                            addr(sr), addr(nl), addr(Ks), addr(betas), addr(fds), addr(its), cv_th, flags,
                                            addr(Tp), addr(mup), addr(epsp), addr(pkp), addr(res))
    _check(rc, "nemb200_nem_device_batch")
    runs = []
    for i, (T, mu, eps, pk) in enumerate(outs):
        r = res[i]
        if want_params:
            mu, eps, pk = mu.cpu().numpy(), eps.cpu().numpy(), pk.cpu().numpy()
        runs.append(NemRun(T, mu, eps, pk, list(r.crit), r.n_iter, bool(r.converged), r.status, r.n_levels, r.elapsed_ms, r.em_ms))
    return runs


def float_chain(d: np.ndarray, g: np.ndarray, parallel: bool):
    """Diagnostics: float32 sums of d and g accumulated from 0 in index order on the device, through the sequential chain
    kernel or the exact parallel scheme (``nemb200_selftest_float_chain``)."""
    _require_cuda()
    d = np.ascontiguousarray(d, np.float32); g = np.ascontiguousarray(g, np.float32)
    assert d.shape == g.shape and d.ndim == 1
    out = (ctypes.c_double * 2)()
    _check(lib().nemb200_selftest_float_chain(d.ctypes.data, g.ctypes.data, d.shape[0], int(parallel), ctypes.cast(out, _vp)),
           "selftest_float_chain")
    return np.float32(out[0]), np.float32(out[1])


def double_chain(l: np.ndarray, z: np.ndarray, parallel: bool):
    """Diagnostics: float32 accumulators receiving double terms (the L / Z criteria), sequential chain or exact parallel scheme."""
    _require_cuda()
    l = np.ascontiguousarray(l, np.float64); z = np.ascontiguousarray(z, np.float64)
    assert l.shape == z.shape and l.ndim == 1
    out = (ctypes.c_double * 2)()
    _check(lib().nemb200_selftest_double_chain(l.ctypes.data, z.ctypes.data, l.shape[0], int(parallel), ctypes.cast(out, _vp)),
           "selftest_double_chain")
    return np.float32(out[0]), np.float32(out[1])


def labels_device(T):
    """(labels int32 tensor with -1 for "S_", entropy) from ``%5.3f``-rounded posteriors (partition.py:206-231)."""
    torch = _require_cuda()
    N, K = T.shape
    lab = torch.empty((N,), dtype=torch.int32, device=T.device)
    ent = ctypes.c_double(0.0)
    with torch.cuda.device(T.device):
        rc = lib().nemb200_labels(_ptr(T), N, K, _ptr(lab), ctypes.byref(ent), _stream())
    _check(rc, "nemb200_labels")
    return lab, float(ent.value)


class Plan:
    """Stepwise interface (one rank's rows): used by the row-sharded multi-GPU driver and the kernel parity tests."""

    def __init__(self, N_local: int, N_total: int, D: int, Ws: int, K: int, free_dispersion: bool = False, flags: int = EPI_REF):
        torch = _require_cuda()
        self.N_local, self.N_total, self.D, self.Ws, self.K = N_local, N_total, D, Ws, K
        self.free_dispersion, self.flags = bool(free_dispersion), flags
        h = _vp()
        _check(lib().nemb200_plan_create(ctypes.byref(h), N_local, N_total, D, Ws, K, int(free_dispersion), flags), "plan_create")
        self._h = h
        cnt, fsum, nkc, nkf, dpad = _vp(), _vp(), _vp(), _vp(), _i64()
        _check(lib().nemb200_stats_ptrs(self._h, ctypes.byref(cnt), ctypes.byref(fsum), ctypes.byref(nkc),
                                        ctypes.byref(nkf), ctypes.byref(dpad)), "stats_ptrs")
        self.Dpad = int(dpad.value)
        self._stat_ptrs = (cnt.value, fsum.value, nkc.value, nkf.value)
        self._views = None

    def close(self):
        if self._h is not None:
            lib().nemb200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats_tensors(self):
        """Torch views (no copy) of the plan's statistic buffers, for in-place all-reduce: cnt int32 [K, Dpad],
        fsum float64 [K, Dpad], nk_cnt int32 [K], nk_fz float64 [K]."""
        if self._views is None:
            torch = _torch()
            K, Dp = self.K, self.Dpad
            dev = torch.cuda.current_device()

            def view(ptr, shape, typestr):
                class _A:
                    pass
                a = _A()
                a.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}
                return torch.as_tensor(a, device=f"cuda:{dev}")
            cnt, fsum, nkc, nkf = self._stat_ptrs
            self._views = (view(cnt, (K, Dp), "<i4"), view(fsum, (K, Dp), "<f8"), view(nkc, (K,), "<i4"), view(nkf, (K,), "<f8"))
        return self._views

    def timing(self, enable: bool):
        """(total ms, launches) of k_mstep_onehot recorded since the last call; then switch event timing on/off."""
        ms, n = _f32(0.0), _i32(0)
        _check(lib().nemb200_plan_timing(self._h, int(enable), ctypes.byref(ms), ctypes.byref(n)), "plan_timing")
        return float(ms.value), int(n.value)

    # ---- row-sharded exchange over NVLink peer memory ----------------------------------------------------------
    def peer_export(self):
        """(64-byte CUDA IPC handle of the plan's block, 4 offsets) to be all-gathered by the driver."""
        h = (ctypes.c_ubyte * 64)()
        offs = (ctypes.c_int64 * PEER_OFFSETS)()
        _check(lib().nemb200_peer_export(self._h, ctypes.cast(h, _vp), ctypes.cast(offs, _vp)), "peer_export")
        return bytes(h), list(offs)

    def peer_attach(self, world: int, rank: int, handles, offsets):
        hb = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        ob = (ctypes.c_int64 * (PEER_OFFSETS * world))(*[v for o in offsets for v in o])
        _check(lib().nemb200_peer_attach(self._h, world, rank, ctypes.cast(hb, _vp), ctypes.cast(ob, _vp)), "peer_attach")

    def peer_logf(self):
        """Torch view of the plan-owned full-length log-density buffer (N_total, K) float64."""
        torch = _torch()
        p = _vp()
        _check(lib().nemb200_peer_logf(self._h, ctypes.byref(p)), "peer_logf")

        class _A:
            pass
        a = _A()
        a.__cuda_array_interface__ = {"shape": (self.N_total, self.K), "typestr": "<f8", "data": (p.value, False), "version": 3}
        return torch.as_tensor(a, device=f"cuda:{torch.cuda.current_device()}")

    def nem_sharded(self, bits_local, row_offset: int, inp: Optional[DeviceInput], beta: float, itermax: int, cv_th: float, T_out):
        """This rank's part of a whole row-sharded run with the device-resident loop (``nemb200_nem_sharded``)."""
        g = inp is not None and inp.rowptr is not None
        res = _Result()
        _check(lib().nemb200_nem_sharded(self._h, _ptr(bits_local), row_offset, _ptr(inp.rowptr) if g else 0, _ptr(inp.col) if g else 0,
                                         _ptr(inp.w) if g else 0, _ptr(inp.sched_ptr) if g else 0, _ptr(inp.sched_rows) if g else 0,
                                         inp.n_levels if g else 1, float(np.float32(beta)), itermax, cv_th, _ptr(T_out),
                                         ctypes.byref(res), _stream()), "nem_sharded")
        return res

    def density_peers(self, bits, row_offset: int):
        _check(lib().nemb200_density_peers(self._h, _ptr(bits), row_offset, _stream()), "density_peers")

    def peer_wait_densities(self):
        _check(lib().nemb200_peer_wait_densities(self._h, _stream()), "peer_wait_densities")

    def stats_blocks(self):
        """(int32 block, fp64 block): flat torch views covering all M-step statistics, for two in-place all-reduces."""
        if getattr(self, "_blocks", None) is None:
            torch = _torch()
            ib, db, il, dl = _vp(), _vp(), _i64(), _i64()
            _check(lib().nemb200_stats_blocks(self._h, ctypes.byref(ib), ctypes.byref(il), ctypes.byref(db), ctypes.byref(dl)), "stats_blocks")
            dev = torch.cuda.current_device()

            def view(ptr, n, typestr):
                class _A:
                    pass
                a = _A()
                a.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}
                return torch.as_tensor(a, device=f"cuda:{dev}")
            self._blocks = (view(ib.value, int(il.value), "<i4"), view(db.value, int(dl.value), "<f8"))
        return self._blocks

    def init_params(self):
        _check(lib().nemb200_init_params(self._h, _stream()), "init_params")

    def mstep_accumulate(self, bits, T):
        _check(lib().nemb200_mstep_accumulate(self._h, _ptr(bits), _ptr(T), _stream()), "mstep_accumulate")

    def mstep_finalize(self):
        _check(lib().nemb200_mstep_finalize(self._h, _stream()), "mstep_finalize")

    def density(self, bits, logf):
        _check(lib().nemb200_density(self._h, _ptr(bits), _ptr(logf), _stream()), "density")

    def sweep(self, N, logf, T_old, T_new, inp: Optional[DeviceInput], beta_now: float, maxdiff):
        g = inp is not None and inp.rowptr is not None
        if g:   # the plan caches a schedule-ordered copy of the graph, keyed by the array addresses: tell it when the
            # tensors behind them are other objects or were written to since the last sweep
            key = tuple((id(t), t._version) for t in (inp.rowptr, inp.col, inp.w, inp.sched_rows))
            if getattr(self, "_graph_key", None) not in (None, key):
                _check(lib().nemb200_plan_graph_changed(self._h), "plan_graph_changed")
            self._graph_key = key
        _check(lib().nemb200_sweep(self._h, N, _ptr(logf), _ptr(T_old), _ptr(T_new), _ptr(inp.rowptr) if g else 0,
                                   _ptr(inp.col) if g else 0, _ptr(inp.w) if g else 0, _ptr(inp.sched_ptr) if g else 0,
                                   _ptr(inp.sched_rows) if g else 0, inp.n_levels if g else 1,
                                   float(np.float32(beta_now)), _ptr(maxdiff) if maxdiff is not None else 0, _stream()), "sweep")

    def read_flags(self):
        """(max |T_new - T_old| of the last sweep that used the plan's own convergence word, status): one 8-byte read."""
        md, st = _f32(0.0), _i32(0)
        _check(lib().nemb200_read_flags(self._h, ctypes.byref(md), ctypes.byref(st), _stream()), "read_flags")
        return float(md.value), int(st.value)

    def criteria(self, N, logf, T, inp: Optional[DeviceInput], beta: float):
        g = inp is not None and inp.rowptr is not None
        out = (ctypes.c_double * 5)()
        _check(lib().nemb200_criteria(self._h, N, _ptr(logf), _ptr(T), _ptr(inp.rowptr) if g else 0, _ptr(inp.col) if g else 0,
                                      _ptr(inp.w) if g else 0, float(np.float32(beta)), ctypes.byref(out), _stream()), "criteria")
        return list(out)

    def get_params_status(self) -> int:
        st = _i32(0)
        _check(lib().nemb200_get_params(self._h, 0, 0, 0, ctypes.byref(st), _stream()), "get_params")
        return int(st.value)

    def get_params(self):
        mu = np.empty((self.K, self.D), np.float32); eps = np.empty((self.K, self.D), np.float32)
        pk = np.empty(self.K, np.float32); st = _i32(0)
        _check(lib().nemb200_get_params(self._h, mu.ctypes.data, eps.ctypes.data, pk.ctypes.data, ctypes.byref(st), _stream()), "get_params")
        return mu, eps, pk, int(st.value)
