"""world_size-2 (and 3) gloo runs of the multi-GPU combine: each rank fills only the segments
it owns, one sum-allreduce makes every rank hold all 64, the fixed final tree gives bits that
do not depend on the world size.  The oracle stands in for the GPU kernel (CPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_paths, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    import oracle
    pkg = entry.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunk = pkg.EUROPEAN_CHUNK
        n_chunks = (n_paths + chunk - 1) // chunk
        seg_lo, seg_hi, c_lo, c_hi = pkg.segment_span(rank, world, n_chunks)
        o = oracle.option(N_PATHS=n_paths)
        partials = np.zeros((n_chunks, 2), dtype=np.float32)
        for c in range(c_lo, c_hi):
            first = c * chunk
            cnt = min(chunk, n_paths - first)
            _, _, pay = oracle.european(o, first, cnt, 1234, oracle.CALL, want_payoffs=True)
            buf = np.zeros(chunk, dtype=np.float32)
            buf[:cnt] = pay
            partials[c] = oracle.chunk_tree_f32(buf, cnt, pkg.EUROPEAN_PATHS_PER_SLOT)
        seg = oracle.segment_tree_f64(partials)
        mine = np.zeros_like(seg)
        mine[seg_lo:seg_hi] = seg[seg_lo:seg_hi]       # +0.0 outside the owned range
        t = torch.from_numpy(mine)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        s, q = oracle.final_tree_f64(t.numpy())
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([s, q]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_combine_is_bit_identical(tmp_path, world, orc, pkg):
    n_paths = 5 * pkg.EUROPEAN_CHUNK + 1234   # ragged tail, 6 chunks
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, n_paths, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    # single-rank truth
    o = orc.option(N_PATHS=n_paths)
    chunk = pkg.EUROPEAN_CHUNK
    partials = []
    for c in range(6):
        first = c * chunk
        cnt = min(chunk, n_paths - first)
        _, _, pay = orc.european(o, first, cnt, 1234, orc.CALL, want_payoffs=True)
        buf = np.zeros(chunk, dtype=np.float32)
        buf[:cnt] = pay
        partials.append(orc.chunk_tree_f32(buf, cnt, pkg.EUROPEAN_PATHS_PER_SLOT))
    seg = orc.segment_tree_f64(np.array(partials, dtype=np.float32))
    s, q = orc.final_tree_f64(seg)
    for g in got:
        assert g[0] == s and g[1] == q      # bit-identical on every rank, for every world size
    ds, _ = orc.european(o, 0, n_paths, 1234, orc.CALL)
    assert s == pytest.approx(ds, rel=1e-6)


def _slab_worker(rank, world, port, n_paths, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    import oracle
    pkg = entry.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = pkg.path_span(rank, world, n_paths)
        o = oracle.option(N_STEPS=16, N_PATHS=n_paths, B=120.0)
        prices, counts = oracle.trajectories(o, lo, hi - lo, 1234)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, prices, counts))   # convenience gather; the data path needs none
        if rank == 0:
            gathered.sort(key=lambda t: t[0])
            np.save(os.path.join(out_dir, "prices.npy"), np.concatenate([g[1] for g in gathered]))
            np.save(os.path.join(out_dir, "counts.npy"), np.concatenate([g[2] for g in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_trajectory_slabs_concatenate_to_the_single_rank_result(tmp_path, world, orc):
    """Trajectory mode / nested MC shard by contiguous path slabs with NO collective: row p depends on
    (seed, p) only, so the ranks' slabs concatenate to exactly the single-rank array."""
    n_paths = 37
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_slab_worker, args=(world, port, n_paths, str(tmp_path)), nprocs=world, join=True)
    prices, counts = orc.trajectories(orc.option(N_STEPS=16, N_PATHS=n_paths, B=120.0), 0, n_paths, 1234)
    assert (np.load(tmp_path / "prices.npy").view(np.uint32) == prices.view(np.uint32)).all()
    assert (np.load(tmp_path / "counts.npy") == counts).all()


class _StandInEngine:
    """The four peer_* methods of Engine without a GPU: records what the host logic asks of it."""

    def __init__(self, pkg, rank, epoch, fail_create=False, fail_connect=False):
        self.pkg, self.rank, self.epoch = pkg, rank, epoch
        self.fail_create, self.fail_connect = fail_create, fail_connect
        self.calls = []

    def peer_epoch(self):
        return self.epoch

    def peer_mailbox_create(self):
        if self.fail_create:
            raise self.pkg.McbError(2, "cudaIpcGetMemHandle failed (stand-in)")
        return bytes([self.rank]) * 64

    def peer_mailbox_connect(self, rank, world, handles, base_epoch):
        if self.fail_connect and world > 1:
            raise self.pkg.McbError(2, "cudaIpcOpenMemHandle failed (stand-in)")
        self.calls.append((rank, world, [h[0] for h in handles], base_epoch))
        self.epoch = max(self.epoch, base_epoch)


def _connect_worker(rank, world, port, scenario, out_dir):
    sys.path.insert(0, ROOT)
    import importlib
    import json
    import __graft_entry__ as entry
    pkg = entry.load_package()
    sharded = importlib.import_module(entry.PKG_NAME + ".sharded")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _StandInEngine(pkg, rank, epoch=10 * (rank + 1),
                             fail_create=(scenario == "create" and rank == 1),
                             fail_connect=(scenario == "connect" and rank == 0))
        ok = sharded.connect_peer_mailboxes(eng, dist, None, rank, world, strict=False)
        with open(os.path.join(out_dir, f"c{rank}.json"), "w") as f:
            json.dump({"ok": ok, "calls": eng.calls, "epoch": eng.epoch}, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scenario", ["fine", "create", "connect"])
def test_peer_transport_agreement_over_gloo(tmp_path, scenario, pkg):
    """The host side of the CUDA-IPC job pipeline (sharded.connect_peer_mailboxes) with world 3 over gloo and
    stand-in engines: all ranks connect with every handle in rank order and the MAXIMUM epoch; if any rank fails to
    export or to map a mailbox, EVERY rank falls back to a group of one (so no rank ever waits for a peer that is
    not in its group) and reports False."""
    import json
    world = 3
    port = 33500 + (os.getpid() % 2000) + {"fine": 0, "create": 1, "connect": 2}[scenario]
    mp.spawn(_connect_worker, args=(world, port, scenario, str(tmp_path)), nprocs=world, join=True)
    got = [json.load(open(tmp_path / f"c{r}.json")) for r in range(world)]
    if scenario == "fine":
        for r, g in enumerate(got):
            assert g["ok"] is True and g["calls"] == [[r, world, [0, 1, 2], 30]] and g["epoch"] == 30
    else:
        for r, g in enumerate(got):
            assert g["ok"] is False
            assert g["calls"][-1][:2] == [0, 1]          # back to a group of one, on every rank
            assert all(c[1] == 1 for c in g["calls"]) or scenario == "connect"
