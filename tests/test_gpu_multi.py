"""Multi-GPU, one process per GPU (NCCL all-reduce and the CUDA-IPC job pipeline): needs >= 2 B200s
(`gpurun --gpus 2`); on a 1-GPU box this skips and tests/test_gpu_shards.py (several shards of ONE engine
on one GPU), tests/test_dist_gloo.py and test_sharded_segments_bit_identical_for_any_world cover the same
logic; bench.py --gpus N prints the price bits at every N."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pkg = entry.load_package()
        sharded = importlib.import_module(entry.PKG_NAME + ".sharded")
        eng = pkg.Engine(rank)
        pricer = sharded.ShardedPricer(eng, transport="nccl")
        assert pricer.transport == "nccl"
        n = 37 * pkg.EUROPEAN_CHUNK + 999
        opt = pkg.option(N_PATHS=n)
        res = pricer.price_european(opt, n, 1234, pkg.CALL)
        put = pricer.price_european(opt, n, 1234, pkg.PUT)
        ob = pkg.option(N_STEPS=50, N_PATHS=20000, B=120.0, P1=5, P2=40)
        bul = pricer.price_bullet(ob, 20000, 1234)
        k = np.linspace(80, 120, 5, dtype=np.float32)
        v = np.linspace(0.1, 0.5, 5, dtype=np.float32)
        sw = pricer.price_sweep(opt, k, v, n, 1234, pkg.CALL)
        # trajectory slabs: no collective, each rank its own contiguous rows
        lo, hi = pkg.path_span(rank, world, 1000)
        rows = eng.simulate_trajectories(pkg.option(N_STEPS=64, N_PATHS=1000), lo, hi - lo, 1234)
        np.save(os.path.join(out_dir, f"r{rank}.npy"),
                np.array([res.sum, res.sumsq, res.price, put.sum, bul.sum, bul.sumsq] + [x.sum for x in sw]))
        np.save(os.path.join(out_dir, f"rows{rank}.npy"), rows)
        # the same through the sharded helper, and a nested-MC slab
        lo2, n2, dev_rows = pricer.trajectories_local(pkg.option(N_STEPS=64, N_PATHS=1000), 1000, 1234)
        nm = pkg.option(N_STEPS=12, N_PATHS=20, N_PATHS_INNER=128, B=120.0, P1=1, P2=10)
        lo3, n3, F = pricer.nested_local(nm, 20, 1234, 1235, pkg.DISCOUNT_CORRECT)
        pricer.synchronize()
        assert (lo2, n2) == (lo, hi - lo) and (dev_rows.cpu().numpy().view(np.uint32) == rows.view(np.uint32)).all()
        np.save(os.path.join(out_dir, f"F{rank}.npy"), F.cpu().numpy())
        # the job pipeline over CUDA-IPC mailboxes: each rank's pricing kernel stores its segments into every
        # rank's mailbox over NVLink; ONE pricing launch + one final pass per job, no collective library.
        # Jobs of >= 64 chunks with a ragged tail (every segment has its own last CTA), 7 epochs so that
        # the 4-slot ring, the tickets, the flags and the acks are all reused; a second pricer on the SAME
        # engine re-connects the mailboxes (epochs are agreed at connect time and stay monotonic).
        eng2 = pkg.Engine(rank)
        peer_out = []
        n_big = 200 * pkg.EUROPEAN_CHUNK + 999
        for generation in range(2):
            peer = sharded.ShardedPricer(eng2, transport="peer")
            assert peer.transport == "peer"
            for k in range(7):
                before = eng2.launch_count
                rp = peer.price_european(opt, n_big + k, 1234, pkg.CALL)
                assert eng2.launch_count - before == 2, "one pricing launch + one final pass per job"
                peer_out += [rp.sum, rp.sumsq, rp.price, float(rp.n_paths)]
            before = eng2.launch_count
            small = peer.price_european(opt, 100_000, 1234, pkg.PUT)   # <= 64 chunks: every rank prices it alone
            assert eng2.launch_count - before == 1, "a small job is one launch on every rank, no final pass"
            peer_out += [small.sum, small.sumsq, small.price, float(small.n_paths)]
            # small (unsharded) and sharded jobs interleaved and pipelined: a sharded job whose mailbox slot was last
            # used by a sharded job many epochs ago, after runs of small ones, must not wait for acks nobody sends
            mixed = [1, 100_000, 64 * pkg.EUROPEAN_CHUNK, 5 * pkg.EUROPEAN_CHUNK + 3, 16_384, n_big + 20, 7, 2 * pkg.EUROPEAN_CHUNK,
                     33, n_big + 21, n_big + 22, 999, n_big + 23]
            for lo in range(0, len(mixed), pkg.PIPELINE_DEPTH):
                tk = [eng2.european_submit(opt, m, 1234, pkg.CALL) for m in mixed[lo:lo + pkg.PIPELINE_DEPTH]]
                for t in tk:
                    rp = eng2.european_collect(t)
                    peer_out += [rp.sum, rp.sumsq, float(rp.n_paths)]
            # pipelined: several jobs in flight, collected afterwards
            tickets = [eng2.european_submit(opt, n_big + 10 + k, 1234, pkg.CALL) for k in range(6)]
            for t in tickets:
                rp = eng2.european_collect(t)
                peer_out += [rp.sum, rp.sumsq]
        assert eng2.peer_timeouts() == 0
        np.save(os.path.join(out_dir, f"peer{rank}.npy"), np.array(peer_out))
        dist.barrier()
        eng2.close()
        eng.close()
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_prices_match_single_gpu_bits(tmp_path, pkg, engine):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n = 37 * pkg.EUROPEAN_CHUNK + 999
    opt = pkg.option(N_PATHS=n)
    res = engine.price_european(opt, n, 1234, pkg.CALL)
    put = engine.price_european(opt, n, 1234, pkg.PUT)
    ob = pkg.option(N_STEPS=50, N_PATHS=20000, B=120.0, P1=5, P2=40)
    bul = engine.price_bullet(ob, 20000, 1234)
    k = np.linspace(80, 120, 5, dtype=np.float32)
    v = np.linspace(0.1, 0.5, 5, dtype=np.float32)
    sw = engine.price_sweep(opt, k, v, n, 1234, pkg.CALL)
    want = np.array([res.sum, res.sumsq, res.price, put.sum, bul.sum, bul.sumsq] + [x.sum for x in sw])
    rows = engine.simulate_trajectories(pkg.option(N_STEPS=64, N_PATHS=1000), 0, 1000, 1234)
    got_rows = []
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert (got == want).all(), (r, got, want)          # bit-identical on every rank
        got_rows.append(np.load(tmp_path / f"rows{r}.npy"))
    assert (np.concatenate(got_rows).view(np.uint32) == rows.view(np.uint32)).all()
    want_peer = []
    n_big = 200 * pkg.EUROPEAN_CHUNK + 999
    for _generation in range(2):
        for k in range(7):
            r1 = engine.price_european(pkg.option(N_PATHS=n_big + k), n_big + k, 1234, pkg.CALL)
            want_peer += [r1.sum, r1.sumsq, r1.price, float(n_big + k)]
        r2 = engine.price_european(pkg.option(), 100_000, 1234, pkg.PUT)
        want_peer += [r2.sum, r2.sumsq, r2.price, 100_000.0]
        for m in [1, 100_000, 64 * pkg.EUROPEAN_CHUNK, 5 * pkg.EUROPEAN_CHUNK + 3, 16_384, n_big + 20, 7, 2 * pkg.EUROPEAN_CHUNK,
                  33, n_big + 21, n_big + 22, 999, n_big + 23]:
            rm = engine.price_european(pkg.option(), m, 1234, pkg.CALL)
            want_peer += [rm.sum, rm.sumsq, float(m)]
        for k in range(6):
            r3 = engine.price_european(pkg.option(), n_big + 10 + k, 1234, pkg.CALL)
            want_peer += [r3.sum, r3.sumsq]
    for r in range(world):
        assert (np.load(tmp_path / f"peer{r}.npy") == np.array(want_peer)).all(), r   # bit-identical, no timeout
    nm = pkg.option(N_STEPS=12, N_PATHS=20, N_PATHS_INNER=128, B=120.0, P1=1, P2=10)
    F, _, _, _ = engine.nested_monte_carlo(nm, 0, 20, 1234, 1235, pkg.DISCOUNT_CORRECT)
    got_F = np.concatenate([np.load(tmp_path / f"F{r}.npy") for r in range(world)])
    assert (got_F.view(np.uint32) == F.view(np.uint32)).all()
