"""Host-side mirror of the reference's call surface over the libmcb200.so C-ABI.

The product is ``libmcb200.so`` (hand-written CUDA for sm_100a + a C++ engine, see
``csrc/`` and ``include/mcb200.h``).  This module is the thin ctypes binding the tests and
``bench.py`` use, and it mirrors the reference's own interface for the path: the
``OptionData`` struct (inc/tool.cuh:13-26) and the ``wrapper_*`` free functions of
inc/wrappers.cuh with the same names, argument meaning and return values (``float`` price,
``-1`` on a launch error).

There is NO CPU fallback: if the CUDA library cannot be loaded or no sm_100 device is present
every entry point raises ``McbError``.  Nothing here imports ``oracle/``.

The directory name contains hyphens, so load it with ``__graft_entry__.load_package()`` (which
registers it as ``monte_carlo_project_cuda_b200``) or ``importlib``.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcb200.so")

SLOTS = 256
SEGMENTS = 64
EUROPEAN_PATHS_PER_SLOT = 64
BULLET_PATHS_PER_SLOT = 4
EUROPEAN_CHUNK = SLOTS * EUROPEAN_PATHS_PER_SLOT
BULLET_CHUNK = SLOTS * BULLET_PATHS_PER_SLOT

CALL, PUT = 0, 1
DISCOUNT_COMPAT, DISCOUNT_CORRECT = 0, 1
HOST, DEVICE = 0, 1

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM, ERR_TIMEOUT = 0, 1, 2, 3, 4, 5
PIPELINE_DEPTH, RESULT_RING, MAX_PEERS = 4, 8, 16
KERNEL_EUROPEAN, KERNEL_BULLET, KERNEL_TRAJECTORY, KERNEL_NESTED, KERNEL_SWEEP, KERNEL_EUROPEAN_PACKED = 0, 1, 2, 3, 4, 5


class McbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"mcb200 status {status}: {message}")
        self.status = status


class OptionData(C.Structure):
    """Byte-compatible with the reference's ``struct OptionData`` (inc/tool.cuh:13-26)."""

    _fields_ = [
        ("S0", C.c_float), ("T", C.c_float), ("K", C.c_float), ("r", C.c_float),
        ("v", C.c_float), ("B", C.c_float),
        ("P1", C.c_int), ("P2", C.c_int), ("N_PATHS", C.c_int), ("N_PATHS_INNER", C.c_int),
        ("N_STEPS", C.c_int), ("step", C.c_float),
    ]

    def __repr__(self):
        return "OptionData(" + ", ".join(f"{n}={getattr(self, n)}" for n, _ in self._fields_) + ")"


def option(S0=100.0, T=1.0, K=100.0, r=0.05, v=0.2, B=120.0, P1=10, P2=50, N_PATHS=1 << 20,
           N_PATHS_INNER=1000, N_STEPS=1, step=None) -> OptionData:
    """BASELINE config-1 parameters by default; ``step`` = T/N_STEPS in float as hello.cu:17."""
    if step is None:   # N_STEPS <= 0 is an invalid option the engine rejects: leave step at 0 instead of dividing
        step = float(np.float32(T) / np.float32(N_STEPS)) if N_STEPS > 0 else 0.0
    return OptionData(S0, T, K, r, v, B, P1, P2, N_PATHS, N_PATHS_INNER, N_STEPS, step)


class Result(C.Structure):
    _fields_ = [("price", C.c_double), ("std_error", C.c_double), ("sum", C.c_double),
                ("sumsq", C.c_double), ("n_paths", C.c_uint64)]

    def __repr__(self):
        return (f"Result(price={self.price:.9g}, std_error={self.std_error:.3g}, sum={self.sum:.12g}, "
                f"sumsq={self.sumsq:.12g}, n_paths={self.n_paths})")


class DeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("sm_count", C.c_int), ("cc_major", C.c_int),
                ("cc_minor", C.c_int), ("clock_khz", C.c_int), ("total_mem", C.c_size_t)]


_u64 = C.c_uint64
_vp = C.c_void_p
_OP = C.POINTER(OptionData)
_RP = C.POINTER(Result)

# name -> (restype, argtypes): every symbol include/mcb200.h declares
SIGNATURES = {
    "mcb_engine_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "mcb_engine_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "mcb_engine_shard_count": (C.c_int, [_vp]),
    "mcb_engine_destroy": (C.c_int, [_vp]),
    "mcb_last_error": (C.c_char_p, []),
    "mcb_version": (C.c_int, []),
    "mcb_get_device_info": (C.c_int, [_vp, C.POINTER(DeviceInfo)]),
    "mcb_synchronize": (C.c_int, [_vp]),
    "mcb_price_european": (C.c_int, [_vp, _OP, _u64, _u64, C.c_int, _RP]),
    "mcb_price_european_packed": (C.c_int, [_vp, _OP, _u64, _u64, C.c_int, _RP]),
    "mcb_european_packed_payoffs": (C.c_int, [_vp, _OP, _u64, _u64, _u64, C.c_int, _vp]),
    "mcb_price_bullet": (C.c_int, [_vp, _OP, _u64, _u64, C.c_int, C.c_float, C.c_int, _RP]),
    "mcb_simulate_trajectories": (C.c_int, [_vp, _OP, _u64, _u64, _u64, _vp, _vp, C.c_int]),
    "mcb_nested_monte_carlo": (C.c_int, [_vp, _OP, _u64, _u64, _u64, _u64, C.c_int, _vp, _vp, _vp, C.c_int,
                                         C.POINTER(C.c_double)]),
    "mcb_price_sweep": (C.c_int, [_vp, _OP, _vp, _vp, C.c_int, _u64, _u64, C.c_int, _RP]),
    "mcb_reduce_sum": (C.c_int, [_vp, _vp, _u64, C.c_int, C.POINTER(C.c_float)]),
    "mcb_price_from_normals": (C.c_int, [_vp, _OP, _vp, _u64, C.c_int, _vp, C.c_int]),
    "mcb_reduce_blocks": (C.c_int, [_vp, _vp, _u64, C.c_int, C.c_uint32, _u64, C.c_int, _vp]),
    "mcb_generate_normals": (C.c_int, [_vp, _u64, _u64, _vp, C.c_int]),
    "mcb_write_trajectories_csv": (C.c_int, [C.c_char_p, _vp, _u64, C.c_int, C.c_float, C.c_float]),
    "mcb_european_segments_async": (C.c_int, [_vp, _OP, _u64, _u64, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "mcb_bullet_segments_async": (C.c_int, [_vp, _OP, _u64, _u64, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                            _vp, _vp]),
    "mcb_sweep_segments_async": (C.c_int, [_vp, _OP, _vp, _vp, C.c_int, _u64, _u64, C.c_int, C.c_int, C.c_int,
                                           _vp, _vp]),
    "mcb_combine_segments_async": (C.c_int, [_vp, _vp, C.c_int, _u64, C.c_float, C.c_float, _vp, _vp]),
    "mcb_trajectories_async": (C.c_int, [_vp, _OP, _u64, _u64, _u64, _vp, _vp, _vp]),
    "mcb_nested_async": (C.c_int, [_vp, _OP, _u64, _u64, _u64, _u64, C.c_int, _vp, _vp, _vp, _vp]),
    "mcb_european_submit": (C.c_int, [_vp, _OP, _u64, _u64, C.c_int, C.POINTER(_u64)]),
    "mcb_european_collect": (C.c_int, [_vp, _u64, _RP]),
    "mcb_pipeline_timer_start": (C.c_int, [_vp]),
    "mcb_pipeline_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "mcb_peer_mailbox_create": (C.c_int, [_vp, _vp]),
    "mcb_peer_epoch": (C.c_int, [_vp, C.POINTER(_u64)]),
    "mcb_peer_mailbox_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _u64]),
    "mcb_set_wait_timeout_ms": (C.c_int, [_vp, _u64]),
    "mcb_peer_timeouts": (C.c_int, [_vp, C.POINTER(_u64)]),
    "mcb_launch_count": (_u64, [_vp]),
    "mcb_timing_enable": (C.c_int, [_vp, C.c_int]),
    "mcb_timing_read": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "mcb_philox_blocks": (C.c_int, [_vp, _u64, _vp, _vp, _u64, _vp]),
    "mcb_curand_blocks": (C.c_int, [_vp, _u64, _vp, _vp, _u64, _vp]),
    "mcb_boxmuller_scan": (C.c_int, [_vp, C.c_int, _u64, _u64, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "mcb_stream_normals": (C.c_int, [_vp, _u64, _u64, _u64, _u64, _vp]),
    "mcb_european_payoffs": (C.c_int, [_vp, _OP, _u64, _u64, _u64, C.c_int, _vp]),
    "mcb_european_chunk_partials": (C.c_int, [_vp, _OP, _u64, _u64, C.c_int, _vp, _u64]),
    "mcb_bullet_payoffs": (C.c_int, [_vp, _OP, _u64, _u64, _u64, C.c_int, C.c_float, C.c_int, _vp]),
    "mcb_last_segments": (C.c_int, [_vp, _vp]),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen libmcb200.so and bind every C-ABI symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise McbError(ERR_NO_DEVICE, f"{path} is not built; run __graft_entry__.build() "
                       "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _check(status):
    if status != OK:
        raise McbError(status, load_library().mcb_last_error().decode("utf-8", "replace"))


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Engine:
    """Persistent handle: streams + workspaces on one GPU (replaces the per-call cudaMalloc /
    cudaFree of every reference wrapper, inc/wrappers.cuh:38-55) -- or, with a list of devices,
    ONE engine over several GPUs of this process (``mcb_engine_create_multi``): the whole-job calls
    then shard internally and return the same bits."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = _vp()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*device)
            _check(self._lib.mcb_engine_create_multi(devs, len(device), C.byref(h)))
            self.devices = list(device)
            self.device = self.devices[0]
        else:
            _check(self._lib.mcb_engine_create(device, C.byref(h)))
            self.device = device
            self.devices = [device]
        self._h = h

    @property
    def shard_count(self) -> int:
        return int(self._lib.mcb_engine_shard_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mcb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- whole-job calls ---------------------------------------------------------------
    def device_info(self) -> DeviceInfo:
        info = DeviceInfo()
        _check(self._lib.mcb_get_device_info(self._h, C.byref(info)))
        return info

    def synchronize(self):
        _check(self._lib.mcb_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.mcb_launch_count(self._h))

    def timing_enable(self, on=True):
        _check(self._lib.mcb_timing_enable(self._h, 1 if on else 0))

    def timing_read(self, kernel):
        """(total device ms, launches) of the recorded launches of ``kernel`` since the last read."""
        ms = C.c_double()
        n = _u64()
        _check(self._lib.mcb_timing_read(self._h, kernel, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def price_european(self, opt, n_paths=0, seed=1234, option_type=CALL) -> Result:
        out = Result()
        _check(self._lib.mcb_price_european(self._h, C.byref(opt), n_paths, seed, option_type, C.byref(out)))
        return out

    def price_european_packed(self, opt, n_paths=0, seed=1234, option_type=CALL) -> Result:
        """Packed keying: path p draws normal p & 3 of subsequence p >> 2 (one Philox block per four paths)."""
        out = Result()
        _check(self._lib.mcb_price_european_packed(self._h, C.byref(opt), n_paths, seed, option_type, C.byref(out)))
        return out

    def european_packed_payoffs(self, opt, first_path, n_paths, seed=1234, option_type=CALL):
        out = np.empty(n_paths, dtype=np.float32)
        _check(self._lib.mcb_european_packed_payoffs(self._h, C.byref(opt), first_path, n_paths, seed, option_type,
                                                     out.ctypes.data))
        return out

    def price_bullet(self, opt, n_paths=0, seed=1234, Ik=0, Sk=0.0, Tk=0) -> Result:
        out = Result()
        _check(self._lib.mcb_price_bullet(self._h, C.byref(opt), n_paths, seed, Ik, Sk, Tk, C.byref(out)))
        return out

    def simulate_trajectories(self, opt, first_path, n_paths, seed=1234, want_counts=False):
        prices = np.empty((n_paths, opt.N_STEPS), dtype=np.float32)
        counts = np.empty((n_paths, opt.N_STEPS), dtype=np.int32) if want_counts else None
        _check(self._lib.mcb_simulate_trajectories(self._h, C.byref(opt), first_path, n_paths, seed,
                                                   prices.ctypes.data, counts.ctypes.data if want_counts else None,
                                                   HOST))
        return (prices, counts) if want_counts else prices

    def trajectories_async(self, opt, first_path, n_paths, seed, d_prices, d_counts=None, stream=None):
        _check(self._lib.mcb_trajectories_async(self._h, C.byref(opt), first_path, n_paths, seed, d_prices, d_counts,
                                                stream))

    def nested_monte_carlo(self, opt, first_outer, n_outer, seed_outer=1234, seed_inner=1235,
                           discount_mode=DISCOUNT_COMPAT):
        F = np.empty((n_outer, opt.N_STEPS), dtype=np.float32)
        prices = np.empty((n_outer, opt.N_STEPS), dtype=np.float32)
        counts = np.empty((n_outer, opt.N_STEPS), dtype=np.int32)
        mean = C.c_double()
        _check(self._lib.mcb_nested_monte_carlo(self._h, C.byref(opt), first_outer, n_outer, seed_outer, seed_inner,
                                                discount_mode, F.ctypes.data, prices.ctypes.data, counts.ctypes.data,
                                                HOST, C.byref(mean)))
        return F, prices, counts, mean.value

    def nested_async(self, opt, first_outer, n_outer, seed_outer, seed_inner, discount_mode, d_F, d_prices=None,
                     d_counts=None, stream=None):
        _check(self._lib.mcb_nested_async(self._h, C.byref(opt), first_outer, n_outer, seed_outer, seed_inner,
                                          discount_mode, d_F, d_prices, d_counts, stream))

    def price_sweep(self, opt, strikes, vols, n_paths=0, seed=1234, option_type=CALL):
        k = _np(strikes, np.float32).ravel()
        v = _np(vols, np.float32).ravel()
        if k.size != v.size:
            raise ValueError("strikes and vols must have the same length")
        out = (Result * k.size)()
        _check(self._lib.mcb_price_sweep(self._h, C.byref(opt), k.ctypes.data, v.ctypes.data, k.size, n_paths, seed,
                                         option_type, out))
        return list(out)

    def reduce_sum(self, x) -> np.float32:
        a = _np(x, np.float32).ravel()
        out = C.c_float()
        _check(self._lib.mcb_reduce_sum(self._h, a.ctypes.data if a.size else None, a.size, HOST, C.byref(out)))
        return np.float32(out.value)

    def reduce_blocks(self, x, n_blocks, span, strided=False):
        """Per-block sums with reduce3..6's index ranges (inc/testing.cuh:185-235)."""
        a = _np(x, np.float32).ravel()
        out = np.empty(n_blocks, dtype=np.float32)
        _check(self._lib.mcb_reduce_blocks(self._h, a.ctypes.data if a.size else None, a.size, HOST, n_blocks, span,
                                           1 if strided else 0, out.ctypes.data))
        return out

    def generate_normals(self, n, seed=1234):
        out = np.empty(n, dtype=np.float32)
        _check(self._lib.mcb_generate_normals(self._h, seed, n, out.ctypes.data if n else None, HOST))
        return out

    def price_from_normals(self, opt, normals):
        z = _np(normals, np.float32)
        n_paths, n_steps = z.shape
        pay = np.empty(n_paths, dtype=np.float32)
        _check(self._lib.mcb_price_from_normals(self._h, C.byref(opt), z.ctypes.data, n_paths, n_steps,
                                                pay.ctypes.data, HOST))
        return pay

    # ---- sharded async pieces (device pointers as ints, stream as int or None) ----------
    def european_segments_async(self, opt, n_paths, seed, option_type, rank, world, d_segments, stream=None):
        _check(self._lib.mcb_european_segments_async(self._h, C.byref(opt), n_paths, seed, option_type, rank, world,
                                                     d_segments, stream))

    def bullet_segments_async(self, opt, n_paths, seed, Ik, Sk, Tk, rank, world, d_segments, stream=None):
        _check(self._lib.mcb_bullet_segments_async(self._h, C.byref(opt), n_paths, seed, Ik, Sk, Tk, rank, world,
                                                   d_segments, stream))

    def sweep_segments_async(self, opt, strikes, vols, n_paths, seed, option_type, rank, world, d_segments,
                             stream=None):
        k = _np(strikes, np.float32).ravel()
        v = _np(vols, np.float32).ravel()
        _check(self._lib.mcb_sweep_segments_async(self._h, C.byref(opt), k.ctypes.data, v.ctypes.data, k.size,
                                                  n_paths, seed, option_type, rank, world, d_segments, stream))

    def combine_segments_async(self, d_segments, n_sets, n_paths, r, T, d_results, stream=None):
        _check(self._lib.mcb_combine_segments_async(self._h, d_segments, n_sets, n_paths, r, T, d_results, stream))

    # ---- the European job pipeline (one launch per shard, result through mapped host memory) --------
    def european_submit(self, opt, n_paths=0, seed=1234, option_type=CALL) -> int:
        t = _u64()
        _check(self._lib.mcb_european_submit(self._h, C.byref(opt), n_paths, seed, option_type, C.byref(t)))
        return int(t.value)

    def european_collect(self, ticket) -> Result:
        out = Result()
        _check(self._lib.mcb_european_collect(self._h, ticket, C.byref(out)))
        return out

    def pipeline_timer_start(self):
        _check(self._lib.mcb_pipeline_timer_start(self._h))

    def pipeline_timer_stop(self) -> float:
        ms = C.c_double()
        _check(self._lib.mcb_pipeline_timer_stop(self._h, C.byref(ms)))
        return ms.value

    # ---- one engine per process: mailboxes mapped over CUDA IPC (NVLink peer stores) ---------------
    def peer_mailbox_create(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self._lib.mcb_peer_mailbox_create(self._h, buf))
        return buf.raw

    def peer_epoch(self) -> int:
        t = _u64()
        _check(self._lib.mcb_peer_epoch(self._h, C.byref(t)))
        return int(t.value)

    def peer_mailbox_connect(self, rank, world, handles, base_epoch):
        blob = b"".join(handles)
        if len(blob) != 64 * world:
            raise ValueError("need one 64-byte handle per rank")
        _check(self._lib.mcb_peer_mailbox_connect(self._h, rank, world, blob, base_epoch))

    def set_wait_timeout_ms(self, ms):
        _check(self._lib.mcb_set_wait_timeout_ms(self._h, int(ms)))

    def peer_timeouts(self) -> int:
        t = _u64()
        _check(self._lib.mcb_peer_timeouts(self._h, C.byref(t)))
        return int(t.value)

    # ---- parity hooks ------------------------------------------------------------------------
    def philox_blocks(self, seed, subsequences, blocks, library=False):
        s = _np(subsequences, np.uint64).ravel()
        b = _np(blocks, np.uint64).ravel()
        out = np.empty((s.size, 4), dtype=np.uint32)
        fn = self._lib.mcb_curand_blocks if library else self._lib.mcb_philox_blocks
        _check(fn(self._h, seed, s.ctypes.data, b.ctypes.data, s.size, out.ctypes.data))
        return out

    def boxmuller_scan(self, which, first_word=0, count=1 << 32):
        """(max |err| vs double, number of bad results) of the radius (0) / sin (1) / cos (2) map."""
        err = C.c_double()
        bad = _u64()
        _check(self._lib.mcb_boxmuller_scan(self._h, which, first_word, count, C.byref(err), C.byref(bad)))
        return err.value, int(bad.value)

    def stream_normals(self, seed, subsequence, count, n0=0):
        out = np.empty(count, dtype=np.float32)
        _check(self._lib.mcb_stream_normals(self._h, seed, subsequence, n0, count, out.ctypes.data))
        return out

    def european_payoffs(self, opt, first_path, n_paths, seed=1234, option_type=CALL):
        out = np.empty(n_paths, dtype=np.float32)
        _check(self._lib.mcb_european_payoffs(self._h, C.byref(opt), first_path, n_paths, seed, option_type,
                                              out.ctypes.data))
        return out

    def european_chunk_partials(self, opt, n_paths, seed=1234, option_type=CALL):
        n_chunks = (n_paths + EUROPEAN_CHUNK - 1) // EUROPEAN_CHUNK
        out = np.empty((n_chunks, 2), dtype=np.float32)
        _check(self._lib.mcb_european_chunk_partials(self._h, C.byref(opt), n_paths, seed, option_type,
                                                     out.ctypes.data, n_chunks))
        return out

    def bullet_payoffs(self, opt, first_path, n_paths, seed=1234, Ik=0, Sk=0.0, Tk=0):
        out = np.empty(n_paths, dtype=np.float32)
        _check(self._lib.mcb_bullet_payoffs(self._h, C.byref(opt), first_path, n_paths, seed, Ik, Sk, Tk,
                                            out.ctypes.data))
        return out

    def last_segments(self):
        out = np.empty((SEGMENTS, 2), dtype=np.float64)
        _check(self._lib.mcb_last_segments(self._h, out.ctypes.data))
        return out


def write_trajectories_csv(path, prices, x0, dt):
    """testing.cu:37-47: ``time,trajectory,value`` with the t = 0 row injected per trajectory."""
    a = _np(prices, np.float32)
    n_traj, n_steps = a.shape
    _check(load_library().mcb_write_trajectories_csv(os.fsencode(path), a.ctypes.data, n_traj, n_steps, x0, dt))


# ---- shard arithmetic (pure host logic, mirrored from csrc/mcb200.cu segment_span) -----------
def segment_span(rank: int, world: int, n_chunks: int):
    """Segments and chunks owned by ``rank`` of ``world``: (seg_lo, seg_hi, chunk_lo, chunk_hi)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    seg_lo = rank * SEGMENTS // world
    seg_hi = (rank + 1) * SEGMENTS // world
    return seg_lo, seg_hi, n_chunks * seg_lo // SEGMENTS, n_chunks * seg_hi // SEGMENTS


def path_span(rank: int, world: int, n_paths: int):
    """Contiguous slab of paths (trajectory mode / NMC outer paths) owned by ``rank``."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return n_paths * rank // world, n_paths * (rank + 1) // world


# ---- the reference's wrapper_* call surface (inc/wrappers.cuh), same names and meaning ------
_default_engine = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def _say(text, quiet):
    if not quiet:
        sys.stdout.write(text)


def wrapper_gpu_option_vanilla(option_data, threadsPerBlock=1024, quiet=False) -> float:
    """inc/wrappers.cuh:33-57.  ``threadsPerBlock`` is accepted and ignored: results do not depend
    on launch geometry.  Seed 1234 as inc/wrappers.cuh:41."""
    try:
        res = default_engine().price_european(option_data, 0, 1234, CALL)
    except McbError:
        return -1.0
    price = float(np.float32(res.price))
    _say(f"Average GPU : {price:g}\n\n", quiet)
    return price


def wrapper_gpu_bullet_option(option_data, threadsPerBlock=1024, quiet=False) -> float:
    """inc/wrappers.cuh:59-93."""
    try:
        res = default_engine().price_bullet(option_data, 0, 1234)
    except McbError:
        return -1.0
    price = float(np.float32(res.price))
    _say(f"Average GPU bullet option : {price:g}\n\n", quiet)
    return price


def wrapper_gpu_bullet_option_atomic(option_data, threadsPerBlock=1024, quiet=False) -> float:
    """inc/wrappers.cuh:95-125.  Same estimator as the non-atomic wrapper; the engine has no
    atomics, so the two are bit-identical here."""
    try:
        res = default_engine().price_bullet(option_data, 0, 1234)
    except McbError:
        return -1.0
    price = float(np.float32(res.price))
    _say(f"Average GPU bullet option atomic : {price:g}\n\n", quiet)
    return price


def _nmc(option_data, label, quiet):
    try:
        _, _, _, mean = default_engine().nested_monte_carlo(option_data, 0, option_data.N_PATHS, 1234, 1235,
                                                            DISCOUNT_COMPAT)
    except McbError:
        return -1.0
    value = float(np.float32(mean))
    _say(f"Average GPU bullet option nmc {label} : {value:g}\n\n", quiet)
    return value


def wrapper_gpu_bullet_option_nmc_one_point_one_block(option_data, threadsPerBlock=1024, number_of_blocks=5000,
                                                      quiet=False) -> float:
    """inc/wrappers.cuh:128-206 (returns the mean-over-points diagnostic, :185-189)."""
    return _nmc(option_data, "one point per block", quiet)


def wrapper_gpu_bullet_option_nmc_one_kernel(option_data, threadsPerBlock=1024, number_of_blocks=5000,
                                             quiet=False) -> float:
    """inc/wrappers.cuh:209-266."""
    return _nmc(option_data, "one kernel", quiet)


def wrapper_gpu_bullet_option_nmc_optimal(option_data, threadsPerBlock=1024, number_of_blocks=5000,
                                          quiet=False) -> float:
    """inc/wrappers.cuh:268-340."""
    return _nmc(option_data, "optimal", quiet)
