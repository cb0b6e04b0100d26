"""ctypes doors onto the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``monte-carlo-project-cuda_b200``) never does: it fails loudly without its CUDA library.

Three shared objects, all built by ``oracle/Makefile``:

* ``libmc_oracle.so``        -- plain-C restatement (``mc_oracle.c``), always buildable (gcc).
* ``_ref/libref_cpu.so``     -- the UNMODIFIED reference CPU pricers / closed form compiled
  from ``/root/reference/inc`` where they lie (present in the build container only; the
  prebuilt file travels to the GPU box).
* ``_ref/libcurand_host.so`` -- cuRAND's own Philox4x32-10 header compiled for the host.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("MCB_REFERENCE", "/root/reference")

SLOTS = 256
SEGMENTS = 64
CALL, PUT = 0, 1
DISCOUNT_COMPAT, DISCOUNT_CORRECT = 0, 1


class OptionData(C.Structure):
    """Same 48-byte layout as the reference's ``struct OptionData`` (inc/tool.cuh:13-26)."""

    _fields_ = [
        ("S0", C.c_float), ("T", C.c_float), ("K", C.c_float), ("r", C.c_float),
        ("v", C.c_float), ("B", C.c_float),
        ("P1", C.c_int), ("P2", C.c_int), ("N_PATHS", C.c_int), ("N_PATHS_INNER", C.c_int),
        ("N_STEPS", C.c_int), ("step", C.c_float),
    ]


def option(S0=100.0, T=1.0, K=100.0, r=0.05, v=0.2, B=120.0, P1=10, P2=50, N_PATHS=1 << 20,
           N_PATHS_INNER=1000, N_STEPS=1, step=None):
    """Config-1 defaults (BASELINE.json configs[0]); ``step`` defaults to T/N_STEPS (hello.cu:17)."""
    if step is None:
        step = float(np.float32(T) / np.float32(N_STEPS))
    return OptionData(S0, T, K, r, v, B, P1, P2, N_PATHS, N_PATHS_INNER, N_STEPS, step)


def build(ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir(os.path.join(REFERENCE, "inc")):
        subprocess.run(["make", "-s", "-C", HERE, "ref", f"REFERENCE={REFERENCE}"], check=True)
    elif ref and not os.path.exists(os.path.join(HERE, "_ref", "libcurand_host.so")):
        # cuRAND's headers ship with the toolkit, so this half of _ref builds anywhere nvcc is.
        subprocess.run(["make", "-s", "-C", HERE, os.path.join(HERE, "_ref", "libcurand_host.so")],
                       check=False)


_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)
_u64 = C.c_uint64

_lib = None
_ref = None
_cur = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libmc_oracle.so")
        src = os.path.join(HERE, "mc_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build(ref=False)
        L = C.CDLL(path)
        P = C.POINTER(OptionData)
        L.orc_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.orc_stream_block.argtypes = [_u64, _u64, _u64, _u32p]
        L.orc_uniform_u.argtypes = [C.c_uint32]; L.orc_uniform_u.restype = C.c_float
        L.orc_angle_v.argtypes = [C.c_uint32]; L.orc_angle_v.restype = C.c_float
        L.orc_stream_normal.argtypes = [_u64, _u64, _u64]; L.orc_stream_normal.restype = C.c_double
        L.orc_stream_normals.argtypes = [_u64, _u64, _u64, _u64, _f64p]
        L.orc_european.argtypes = [P, _u64, _u64, _u64, C.c_int, _f64p, _f64p, _f32p]
        L.orc_european_packed.argtypes = [P, _u64, _u64, _u64, C.c_int, _f64p, _f64p, _f32p]
        L.orc_bullet.argtypes = [P, _u64, _u64, _u64, C.c_int, C.c_float, C.c_int, _f64p, _f64p, _f32p]
        L.orc_trajectories.argtypes = [P, _u64, _u64, _u64, _f32p, _i32p]
        L.orc_nmc.argtypes = [P, _u64, _u64, _u64, _u64, C.c_int, _f32p, _f32p, _i32p]
        L.orc_sweep.argtypes = [P, _f32p, _f32p, C.c_int, _u64, _u64, _u64, C.c_int, _f64p, _f64p]
        L.orc_pregen_european.argtypes = [P, _f32p, _u64, C.c_int, _f32p]
        L.orc_price_from_sum.argtypes = [C.c_double, _u64, C.c_float, C.c_float]
        L.orc_price_from_sum.restype = C.c_double
        L.orc_std_error.argtypes = [C.c_double, C.c_double, _u64, C.c_float, C.c_float]
        L.orc_std_error.restype = C.c_double
        L.orc_cnd_reference.argtypes = [C.c_float]; L.orc_cnd_reference.restype = C.c_float
        L.orc_bs_call_reference.argtypes = [C.c_float] * 5; L.orc_bs_call_reference.restype = C.c_float
        L.orc_bs_call_exact.argtypes = [C.c_double] * 5; L.orc_bs_call_exact.restype = C.c_double
        L.orc_bs_put_exact.argtypes = [C.c_double] * 5; L.orc_bs_put_exact.restype = C.c_double
        L.orc_chunk_tree_f32.argtypes = [_f32p, _u64, C.c_int, _f32p, _f32p]
        L.orc_segment_range.argtypes = [_u64, C.c_int, C.POINTER(_u64), C.POINTER(_u64)]
        L.orc_segment_tree_f64.argtypes = [_f32p, _u64, _f64p]
        L.orc_final_tree_f64.argtypes = [_f64p, _f64p, _f64p]
        L.orc_reduce_sum_f32.argtypes = [_f32p, _u64]; L.orc_reduce_sum_f32.restype = C.c_float
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_cpu.so"))


def ref_cpu():
    """The unmodified reference CPU pricers (oracle/_ref/libref_cpu.so)."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libref_cpu.so")
        if not os.path.exists(path):
            build(ref=True)
        L = C.CDLL(path)
        P = C.POINTER(OptionData)
        L.ref_sizeof_option_data.restype = C.c_int
        L.ref_vanilla_cpu.argtypes = [P]; L.ref_vanilla_cpu.restype = C.c_float
        L.ref_bullet_cpu.argtypes = [P]; L.ref_bullet_cpu.restype = C.c_float
        L.ref_vanilla_cpu_chunked.argtypes = [P, _u64, C.c_int, C.POINTER(_u64)]
        L.ref_vanilla_cpu_chunked.restype = C.c_double
        L.ref_bullet_cpu_chunked.argtypes = [P, _u64, C.c_int, C.POINTER(_u64)]
        L.ref_bullet_cpu_chunked.restype = C.c_double
        L.ref_black_scholes.argtypes = [C.c_float] * 5; L.ref_black_scholes.restype = C.c_float
        L.ref_cnd.argtypes = [C.c_float]; L.ref_cnd.restype = C.c_float
        _ref = L
    return _ref


def have_curand_host() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libcurand_host.so"))


def curand_host():
    """cuRAND's own Philox4x32-10 compiled for the host (oracle/_ref/libcurand_host.so)."""
    global _cur
    if _cur is None:
        path = os.path.join(HERE, "_ref", "libcurand_host.so")
        if not os.path.exists(path):
            build(ref=True)
        L = C.CDLL(path)
        L.curand_host_block.argtypes = [_u64, _u64, _u64, _u32p]
        L.curand_host_philox.argtypes = [_u32p, _u32p, _u32p]
        L.curand_host_words.argtypes = [_u64, _u64, _u64, C.c_int, _u32p]
        L.curand_host_normals.argtypes = [_u64, _u64, C.c_int, _f32p]
        _cur = L
    return _cur


# ---- numpy-friendly wrappers ------------------------------------------------------------

def _p(a, t):
    return a.ctypes.data_as(t)


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, _u32p), _p(k, _u32p), _p(out, _u32p))
    return out


def stream_block(seed, subsequence, block):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_stream_block(seed, subsequence, block, _p(out, _u32p))
    return out


def stream_normals(seed, subsequence, count, n0=0):
    out = np.zeros(count, dtype=np.float64)
    lib().orc_stream_normals(seed, subsequence, n0, count, _p(out, _f64p))
    return out


def european(o, first_path, n_paths, seed=1234, option_type=CALL, want_payoffs=False):
    s = C.c_double(); q = C.c_double()
    pay = np.zeros(n_paths, dtype=np.float32) if want_payoffs else None
    lib().orc_european(C.byref(o), first_path, n_paths, seed, option_type, C.byref(s), C.byref(q),
                       _p(pay, _f32p) if want_payoffs else None)
    return (s.value, q.value, pay) if want_payoffs else (s.value, q.value)


def european_packed(o, first_path, n_paths, seed=1234, option_type=CALL, want_payoffs=False):
    """Packed keying: path p draws normal p & 3 of subsequence p >> 2 (orc_european_packed)."""
    s = C.c_double(); q = C.c_double()
    pay = np.zeros(n_paths, dtype=np.float32) if want_payoffs else None
    lib().orc_european_packed(C.byref(o), first_path, n_paths, seed, option_type, C.byref(s), C.byref(q),
                              _p(pay, _f32p) if want_payoffs else None)
    return (s.value, q.value, pay) if want_payoffs else (s.value, q.value)


def bullet(o, first_path, n_paths, seed=1234, Ik=0, Sk=0.0, Tk=0, want_payoffs=False):
    s = C.c_double(); q = C.c_double()
    pay = np.zeros(n_paths, dtype=np.float32) if want_payoffs else None
    lib().orc_bullet(C.byref(o), first_path, n_paths, seed, Ik, Sk, Tk, C.byref(s), C.byref(q),
                     _p(pay, _f32p) if want_payoffs else None)
    return (s.value, q.value, pay) if want_payoffs else (s.value, q.value)


def trajectories(o, first_path, n_paths, seed=1234, want_counts=True):
    prices = np.zeros((n_paths, o.N_STEPS), dtype=np.float32)
    counts = np.zeros((n_paths, o.N_STEPS), dtype=np.int32) if want_counts else None
    lib().orc_trajectories(C.byref(o), first_path, n_paths, seed, _p(prices, _f32p),
                           _p(counts, _i32p) if want_counts else None)
    return prices, counts


def nmc(o, first_outer, n_outer, seed_outer=1234, seed_inner=1235, discount_mode=DISCOUNT_COMPAT):
    F = np.zeros((n_outer, o.N_STEPS), dtype=np.float32)
    prices = np.zeros((n_outer, o.N_STEPS), dtype=np.float32)
    counts = np.zeros((n_outer, o.N_STEPS), dtype=np.int32)
    lib().orc_nmc(C.byref(o), first_outer, n_outer, seed_outer, seed_inner, discount_mode,
                  _p(F, _f32p), _p(prices, _f32p), _p(counts, _i32p))
    return F, prices, counts


def sweep(o, strikes, vols, first_path, n_paths, seed=1234, option_type=CALL):
    k = np.ascontiguousarray(strikes, dtype=np.float32); v = np.ascontiguousarray(vols, dtype=np.float32)
    assert k.shape == v.shape
    s = np.zeros(k.size, dtype=np.float64); q = np.zeros(k.size, dtype=np.float64)
    lib().orc_sweep(C.byref(o), _p(k, _f32p), _p(v, _f32p), k.size, first_path, n_paths, seed,
                    option_type, _p(s, _f64p), _p(q, _f64p))
    return s, q


def pregen_european(o, normals):
    z = np.ascontiguousarray(normals, dtype=np.float32)
    n_paths, n_steps = z.shape
    pay = np.zeros(n_paths, dtype=np.float32)
    lib().orc_pregen_european(C.byref(o), _p(z, _f32p), n_paths, n_steps, _p(pay, _f32p))
    return pay


def price_from_sum(s, n, r, T):
    return lib().orc_price_from_sum(s, n, r, T)


def std_error(s, q, n, r, T):
    return lib().orc_std_error(s, q, n, r, T)


def chunk_tree_f32(payoffs, n_valid, paths_per_slot):
    p = np.ascontiguousarray(payoffs, dtype=np.float32)
    s = C.c_float(); q = C.c_float()
    lib().orc_chunk_tree_f32(_p(p, _f32p), n_valid, paths_per_slot, C.byref(s), C.byref(q))
    return np.float32(s.value), np.float32(q.value)


def segment_range(n_chunks, segment):
    lo = _u64(); hi = _u64()
    lib().orc_segment_range(n_chunks, segment, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def segment_tree_f64(chunk_partials):
    cp = np.ascontiguousarray(chunk_partials, dtype=np.float32).reshape(-1, 2)
    seg = np.zeros((SEGMENTS, 2), dtype=np.float64)
    lib().orc_segment_tree_f64(_p(cp, _f32p), cp.shape[0], _p(seg, _f64p))
    return seg


def final_tree_f64(segments):
    seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(SEGMENTS, 2)
    s = C.c_double(); q = C.c_double()
    lib().orc_final_tree_f64(_p(seg, _f64p), C.byref(s), C.byref(q))
    return s.value, q.value


def reduce_sum_f32(x):
    a = np.ascontiguousarray(x, dtype=np.float32)
    return np.float32(lib().orc_reduce_sum_f32(_p(a, _f32p), a.size))
