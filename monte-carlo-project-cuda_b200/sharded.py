"""Multi-GPU pricing: one process per GPU, paths sharded by index, ONE exchange of 1 KiB.

The reference is single-GPU ("Parallelism strategies: none"); this is the
additive capability north_star asks for.  Every path is a pure function of (seed, path id), so
rank g of G prices the chunks of its own 64/G reduction segments and writes +0.0 into the
segments it does not own; one ``all_reduce(SUM)`` of the 1 KiB segment vector (x + 0.0 is exact,
so the result does not depend on NCCL's ring/tree order) leaves all 64 double segments on every
rank, and the fixed final tree (``mcb_combine_segments_async``) gives a price whose bits do not
depend on G.  ``torch`` / ``torch.distributed`` are plumbing only: device memory, the stream,
the NCCL communicator.  Trajectory mode and nested MC shard by contiguous path slabs with no
collective at all (``path_span``).

Nothing here imports ``oracle/``; without the CUDA library every call raises ``McbError``.
"""
from __future__ import annotations

import ctypes as C

from . import (CALL, SEGMENTS, Engine, McbError, Result, path_span)  # noqa: F401  (re-exported helpers)


def connect_peer_mailboxes(engine, dist, group, rank: int, world: int, strict: bool = False) -> bool:
    """CUDA-IPC mailbox exchange of a process group; EVERY rank ends up in the same mode.

    Each rank exports its mailbox and reports its job epoch; handles and epochs are all-gathered; every rank
    connects with the MAXIMUM epoch (a fresh or lagging engine can then neither satisfy nor block a wait of the
    group's jobs); the outcomes are all-gathered again (that is also the barrier between mapping and the first peer
    store).  If ANY rank failed anywhere, every rank goes back to a group of one (its engine keeps pricing on its
    own) and the function returns False -- the caller then uses the all-reduce route -- or raises when ``strict``.
    Pure host logic over ``dist`` and the engine's four peer_* methods (tests/test_dist_gloo.py drives it on CPU
    with stand-in engines)."""
    mine = {"handle": None, "epoch": engine.peer_epoch(), "error": None}
    try:
        if world > 16:
            raise McbError(1, "the peer transport supports at most 16 ranks")
        mine["handle"] = engine.peer_mailbox_create()
    except McbError as exc:
        mine["error"] = str(exc)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    ok = all(x["error"] is None for x in everyone)
    base = max(x["epoch"] for x in everyone)
    err = None
    if ok:
        try:
            engine.peer_mailbox_connect(rank, world, [x["handle"] for x in everyone], base)
        except McbError as exc:
            err = str(exc)
    results = [None] * world
    dist.all_gather_object(results, err, group=group)   # also the barrier after connect
    ok = ok and all(x is None for x in results)
    if not ok:
        # back to a group of one, so that engine.price_european keeps working on this rank (a group of one maps
        # nothing: its handle slot is never opened, so a rank whose export failed can pass a blank one)
        engine.peer_mailbox_connect(0, 1, [mine["handle"] or bytes(64)], engine.peer_epoch())
        if strict:
            raise McbError(2, "peer transport unavailable: " + "; ".join(
                str(x) for x in ([y["error"] for y in everyone] + results) if x))
    dist.barrier(group=group)
    return ok


class ShardedPricer:
    """European / bullet / sweep pricing over the ranks of a ``torch.distributed`` group.

    ``transport`` says how the 64 (sum, sumsq) segments of a European job cross GPUs:

    * ``"peer"`` -- the engine's job pipeline (``mcb_european_submit`` / ``collect``): every rank's
      pricing kernel stores the segments it owns straight into every rank's mailbox over NVLink
      (CUDA-IPC mapped peer memory) and the final tree runs on the engine's second stream; ONE pricing
      launch per job, no collective library, the pricing stream never waits for a peer;
    * ``"nccl"`` -- segment pass, one ``all_reduce(SUM)`` of the 1 KiB segment vector, final tree,
      all enqueued on ``torch.cuda.current_stream()``;
    * ``"auto"`` (default) -- ``"peer"`` when CUDA IPC connects on EVERY rank, else ``"nccl"``.

    Both give the bits of a single-GPU run.  Bullet and sweep always take the all-reduce path.
    For the all-reduce path the C-ABI reads a NULL stream as "the engine's own stream", so callers must
    make a real (non-default) torch stream current -- ``with torch.cuda.stream(pricer.stream):``.
    The peer path needs at most one rank per GPU and a world of at most 16.
    """

    def __init__(self, engine: Engine, group=None, max_sets: int = 1, transport: str = "auto"):
        import torch
        import torch.distributed as dist

        if transport not in ("auto", "nccl", "peer"):
            raise ValueError("transport must be 'auto', 'nccl' or 'peer'")
        self.torch = torch
        self.dist = dist
        self.engine = engine
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.device = torch.device("cuda", engine.device)
        self.stream = torch.cuda.Stream(self.device)
        with torch.cuda.stream(self.stream):     # the buffers are zeroed on the stream that will use them
            self._reserve(max_sets)
        self.stream.synchronize()
        self._ticket = None
        self.transport = "nccl"
        if self.world == 1:
            self.transport = "peer"              # a group of one: the plain single-launch pipeline
        elif transport in ("auto", "peer"):
            self.transport = "peer" if self._connect_peers(strict=transport == "peer") else "nccl"

    def _connect_peers(self, strict: bool) -> bool:
        return connect_peer_mailboxes(self.engine, self.dist, self.group, self.rank, self.world, strict)

    def _reserve(self, n_sets: int):
        t = self.torch
        self.max_sets = n_sets
        self.segments = t.zeros(n_sets * 2 * SEGMENTS, dtype=t.float64, device=self.device)
        self.results = t.zeros(n_sets * 5, dtype=t.float64, device=self.device)      # mcb_result = 40 bytes
        self.h_results = t.zeros(n_sets * 5, dtype=t.float64).pin_memory()

    def _stream(self):
        ptr = self.torch.cuda.current_stream(self.device).cuda_stream
        if not ptr:
            raise RuntimeError("the legacy default stream is current: wrap the call in "
                               "`with torch.cuda.stream(pricer.stream):`")
        return ptr

    def _finish_async(self, n_sets, n_paths, r, T):
        if self.world > 1:
            self.dist.all_reduce(self.segments[: n_sets * 2 * SEGMENTS], op=self.dist.ReduceOp.SUM, group=self.group)
        self.engine.combine_segments_async(self.segments.data_ptr(), n_sets, n_paths, r, T,
                                           self.results.data_ptr(), self._stream())

    # ---- enqueue-only ---------------------------------------------------------------------
    def european_async(self, opt, n_paths, seed=1234, option_type=CALL):
        """Enqueue one European job; ``european_result()`` fetches the latest one."""
        if self.transport == "peer":
            self._ticket = self.engine.european_submit(opt, n_paths, seed, option_type)
            return
        self._ticket = None
        self.engine.european_segments_async(opt, n_paths, seed, option_type, self.rank, self.world,
                                            self.segments.data_ptr(), self._stream())
        self._finish_async(1, n_paths, opt.r, opt.T)

    def european_result(self) -> Result:
        if self.transport == "peer":
            if self._ticket is None:
                raise RuntimeError("no European job has been submitted")
            return self.engine.european_collect(self._ticket)
        return self._fetch(1)[0]

    def bullet_async(self, opt, n_paths, seed=1234, Ik=0, Sk=0.0, Tk=0):
        self.engine.bullet_segments_async(opt, n_paths, seed, Ik, Sk, Tk, self.rank, self.world,
                                          self.segments.data_ptr(), self._stream())
        self._finish_async(1, n_paths, opt.r, opt.T)

    def sweep_async(self, opt, strikes, vols, n_paths, seed=1234, option_type=CALL):
        n_sets = len(strikes)
        if n_sets > self.max_sets:
            with self.torch.cuda.stream(self.stream):
                self._reserve(n_sets)
            self.stream.synchronize()
        self.engine.sweep_segments_async(opt, strikes, vols, n_paths, seed, option_type, self.rank, self.world,
                                         self.segments.data_ptr(), self._stream())
        self._finish_async(n_sets, n_paths, opt.r, opt.T)

    # ---- synchronous: result on the host ------------------------------------------------
    def _fetch(self, n_sets):
        self.h_results[: n_sets * 5].copy_(self.results[: n_sets * 5], non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        out = (Result * n_sets)()
        C.memmove(out, self.h_results.data_ptr(), C.sizeof(Result) * n_sets)
        for r in out:
            if r.n_paths == 0 or r.price != r.price:
                raise McbError(5, "a sharded result is poisoned (n_paths == 0 / NaN): a rank did not deliver")
        return list(out)

    def price_european(self, opt, n_paths, seed=1234, option_type=CALL) -> Result:
        if self.transport == "peer":
            return self.engine.european_collect(self.engine.european_submit(opt, n_paths, seed, option_type))
        with self.torch.cuda.stream(self.stream):
            self.european_async(opt, n_paths, seed, option_type)
            return self._fetch(1)[0]

    def price_bullet(self, opt, n_paths, seed=1234, Ik=0, Sk=0.0, Tk=0) -> Result:
        with self.torch.cuda.stream(self.stream):
            self.bullet_async(opt, n_paths, seed, Ik, Sk, Tk)
            return self._fetch(1)[0]

    def price_sweep(self, opt, strikes, vols, n_paths, seed=1234, option_type=CALL):
        with self.torch.cuda.stream(self.stream):
            self.sweep_async(opt, strikes, vols, n_paths, seed, option_type)
            return self._fetch(len(strikes))

    # ---- slab-sharded modes: no collective on the data path --------------------------------
    def trajectories_local(self, opt, n_paths, seed=1234, d_prices=None, d_counts=None):
        """This rank's contiguous slab of the n_paths trajectories: returns (first_path, n_local,
        prices tensor [n_local, N_STEPS] on this GPU).  Row p is a pure function of (seed, p), so the
        concatenation over ranks is bit-identical to a single-GPU run."""
        t = self.torch
        lo, hi = path_span(self.rank, self.world, n_paths)
        n_local = hi - lo
        if d_prices is None:
            d_prices = t.empty((max(n_local, 1), opt.N_STEPS), dtype=t.float32, device=self.device)
        if n_local:
            with t.cuda.stream(self.stream):
                self.engine.trajectories_async(opt, lo, n_local, seed, d_prices.data_ptr(),
                                               d_counts.data_ptr() if d_counts is not None else None, self._stream())
        return lo, n_local, d_prices[:n_local]

    def nested_local(self, opt, n_outer, seed_outer=1234, seed_inner=1235, discount_mode=0):
        """This rank's slab of outer trajectories of a nested Monte Carlo: (first_outer, n_local,
        F tensor [n_local, N_STEPS]).  Outer path p and all its inner paths depend on p only."""
        t = self.torch
        lo, hi = path_span(self.rank, self.world, n_outer)
        n_local = hi - lo
        F = t.empty((max(n_local, 1), opt.N_STEPS), dtype=t.float32, device=self.device)
        if n_local:
            with t.cuda.stream(self.stream):
                self.engine.nested_async(opt, lo, n_local, seed_outer, seed_inner, discount_mode, F.data_ptr(),
                                         None, None, self._stream())
        return lo, n_local, F[:n_local]

    def synchronize(self):
        self.stream.synchronize()
