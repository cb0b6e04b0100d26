"""The sharded path through the C-ABI alone (mcb_engine_create_multi, no Python collective): ONE engine
over several shards must return the bits of a single-device engine for every whole-job call.

Runs on a 1-GPU box: a device may be listed several times (`[0, 0, 0]` = three shards on GPU 0), which
drives exactly the code a multi-GPU engine runs -- per-shard launches, segment tickets, peer stores into
the leader's mailbox, the ring of mailbox slots, event pacing, the final pass on the second stream;
only the NVLink hop is missing (tests/test_gpu_multi.py and bench.py --gpus N cover that on real GPUs).
In-process shards never spin on one another (event dependencies), so sharing a GPU is safe.
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_lists():
    import torch
    lists = [[0, 0], [0, 0, 0], [0] * 8]
    n = torch.cuda.device_count()
    if n >= 2:
        lists.append(list(range(min(n, 8))))
        lists.append([0, 1])
    return lists


def _bits(r):
    return (np.float64(r.sum).view(np.uint64), np.float64(r.sumsq).view(np.uint64),
            np.float64(r.price).view(np.uint64), np.float64(r.std_error).view(np.uint64), r.n_paths)


@pytest.fixture(scope="module")
def multi_engines(pkg):
    engines = [pkg.Engine(d) for d in _device_lists()]
    yield engines
    for e in engines:
        e.close()


def test_single_engine_prices_in_one_launch(pkg, engine, orc):
    """mcb_price_european = ONE kernel launch: pricing, segment folds, final tree and the host-visible
    result all come from european_small_job_kernel (jobs of at most 64 chunks: a cluster of eight CTAs per
    chunk) or european_job_kernel (the reference: two launches, a sync and a copy, inc/wrappers.cuh:41-49).
    The sizes walk the cluster kernel's edges: one lane, one warp +- 1, one slot row +- 1, half a chunk, a
    chunk +- 1, ragged multi-chunk jobs, the largest small job and the first job past it."""
    C = pkg.EUROPEAN_CHUNK
    for i, n in enumerate((1, 2, 31, 32, 33, 255, 256, 257, 2047, 2048, 2049, 8191, C - 1, C, C + 1, 3 * C + 5, 100_000,
                           1_000_000, 64 * C - 1, 64 * C, 64 * C + 1, 200 * C + 999)):
        opt = pkg.option(N_PATHS=n)
        typ = pkg.PUT if i % 3 == 2 else pkg.CALL
        before = engine.launch_count
        res = engine.price_european(opt, n, 1234, typ)
        assert engine.launch_count - before == 1
        # the segments the kernel left in mapped host memory fold to the result with the oracle's tree
        seg = engine.last_segments()
        s, q = orc.final_tree_f64(seg)
        assert s == res.sum and q == res.sumsq and res.n_paths == n
        # ... and are the oracle's segment tree of the kernel's own chunk partials
        cp = engine.european_chunk_partials(opt, n, 1234, typ)
        assert (orc.segment_tree_f64(cp) == seg).all(), n


def test_large_shards_take_the_two_launch_route_with_the_same_bits(pkg, engine, multi_engines):
    """A shard of >= 4096 chunks prices with the plain kernel + one segment launch (the per-CTA ticket of
    the fused kernel costs more than a launch there); smaller shards use the single fused launch.  Same
    tree, same bits: the job below is two-launch on one device and fused on two or more shards."""
    n = 4100 * pkg.EUROPEAN_CHUNK + 999
    opt = pkg.option(N_PATHS=1)
    before = engine.launch_count
    want = engine.price_european(opt, n, 1234, pkg.CALL)
    assert engine.launch_count - before == 2
    seg = engine.last_segments()
    cp = engine.european_chunk_partials(opt, n, 1234, pkg.CALL)
    import oracle
    assert (oracle.segment_tree_f64(cp) == seg).all()
    assert oracle.final_tree_f64(seg) == (want.sum, want.sumsq)
    for multi in multi_engines:
        before = multi.launch_count
        got = multi.price_european(opt, n, 1234, pkg.CALL)
        assert multi.launch_count - before == multi.shard_count + 1
        assert _bits(got) == _bits(want)
    big = 2 * 4096 * pkg.EUROPEAN_CHUNK + 5          # two shards of 4096 chunks: two-launch on each
    two = multi_engines[0]
    before = two.launch_count
    got = two.price_european(opt, big, 1234, pkg.PUT)
    assert two.launch_count - before == 2 * 2 + 1
    assert _bits(got) == _bits(engine.price_european(opt, big, 1234, pkg.PUT))


def test_multi_engine_european_bits(pkg, engine, multi_engines):
    """Jobs of at most 64 chunks are priced by the group's leader alone (ONE launch: sharding a 10 us job costs more
    than it saves, and the tree does not depend on the sharding); larger ones by every shard + the final pass."""
    sizes = (1, 100_000,                                # small: the leader alone
             63 * pkg.EUROPEAN_CHUNK + 5, 64 * pkg.EUROPEAN_CHUNK,
             64 * pkg.EUROPEAN_CHUNK + 1,               # 65 chunks: most shards own one or two segments' worth
             200 * pkg.EUROPEAN_CHUNK + 999,            # >= 64 chunks, ragged tail
             1 << 24)
    for multi in multi_engines:
        k = multi.shard_count
        assert k == len(multi.devices)
        for rep in range(3):                            # 3 x 14 jobs, small and sharded interleaved: the ring wraps
            for n in sizes:
                for typ in (pkg.CALL, pkg.PUT):
                    opt = pkg.option(N_PATHS=n, K=100.0 + rep)
                    before = multi.launch_count
                    got = multi.price_european(opt, n, 1234 + rep, typ)
                    small = -(-n // pkg.EUROPEAN_CHUNK) <= 64
                    assert multi.launch_count - before == (1 if small else k + 1)   # one per shard + the final pass
                    want = engine.price_european(opt, n, 1234 + rep, typ)
                    assert _bits(got) == _bits(want), (multi.devices, n, typ, got, want)
        assert multi.peer_timeouts() == 0


def test_pipelined_jobs_and_ticket_rules(pkg, engine, multi_engines):
    for eng in [engine] + multi_engines:
        jobs = [(100_000 + 1_000_003 * i, 1234 + i, pkg.PUT if i & 1 else pkg.CALL) for i in range(pkg.RESULT_RING + 3)]
        tickets = [eng.european_submit(pkg.option(N_PATHS=n), n, seed, typ) for n, seed, typ in jobs]
        assert tickets == list(range(tickets[0], tickets[0] + len(jobs)))
        # only the last RESULT_RING results are kept
        for t in tickets[:3]:
            with pytest.raises(pkg.McbError):
                eng.european_collect(t)
        for (n, seed, typ), t in list(zip(jobs, tickets))[3:]:
            got = eng.european_collect(t)
            want = engine.price_european(pkg.option(N_PATHS=n), n, seed, typ)
            assert _bits(got) == _bits(want)
        with pytest.raises(pkg.McbError):
            eng.european_collect(tickets[-1] + 1_000_000)      # never submitted
        eng.pipeline_timer_start()
        last = [eng.european_submit(pkg.option(), 1 << 22, 1234, pkg.CALL) for _ in range(6)][-1]
        ms = eng.pipeline_timer_stop()
        assert 0.0 < ms < 1000.0
        assert _bits(eng.european_collect(last)) == _bits(engine.price_european(pkg.option(), 1 << 22, 1234, pkg.CALL))


def test_multi_engine_bullet_sweep_trajectories_nested(pkg, engine, multi_engines):
    ob = pkg.option(N_STEPS=50, N_PATHS=20000, B=120.0, P1=5, P2=40)
    k = np.linspace(80, 120, 7, dtype=np.float32)
    v = np.linspace(0.1, 0.5, 7, dtype=np.float32)
    n = 37 * pkg.EUROPEAN_CHUNK + 999
    ot = pkg.option(N_STEPS=252, N_PATHS=1001, B=110.0)
    nm = pkg.option(N_STEPS=12, N_PATHS=21, N_PATHS_INNER=128, B=120.0, P1=1, P2=10)
    want_pk = engine.price_european_packed(pkg.option(), n, 1234, pkg.PUT)
    want_b = engine.price_bullet(ob, 20000, 1234)
    want_s = engine.price_sweep(pkg.option(), k, v, n, 1234, pkg.CALL)
    want_rows, want_counts = engine.simulate_trajectories(ot, 5, 1001, 1234, want_counts=True)
    want_F, want_P, want_C, want_mean = engine.nested_monte_carlo(nm, 0, 21, 1234, 1235, pkg.DISCOUNT_CORRECT)
    for multi in multi_engines:
        assert _bits(multi.price_bullet(ob, 20000, 1234)) == _bits(want_b)
        assert _bits(multi.price_european_packed(pkg.option(), n, 1234, pkg.PUT)) == _bits(want_pk)
        got_s = multi.price_sweep(pkg.option(), k, v, n, 1234, pkg.CALL)
        assert [_bits(r) for r in got_s] == [_bits(r) for r in want_s]
        rows, counts = multi.simulate_trajectories(ot, 5, 1001, 1234, want_counts=True)
        assert (rows.view(np.uint32) == want_rows.view(np.uint32)).all() and (counts == want_counts).all()
        F, P, Cn, mean = multi.nested_monte_carlo(nm, 0, 21, 1234, 1235, pkg.DISCOUNT_CORRECT)
        assert (F.view(np.uint32) == want_F.view(np.uint32)).all() and (Cn == want_C).all()
        assert (P.view(np.uint32) == want_P.view(np.uint32)).all() and mean == want_mean


def test_c_program_on_a_multi_device_engine(pkg):
    """examples/price_c.c (C99, no Python, no CUDA headers) on ONE engine over several shards prints the
    bits of the single-device run: multi-GPU is reachable from the C-ABI alone."""
    import torch
    import __graft_entry__ as entry
    entry.build_examples()
    exe = os.path.join(ROOT, "build", "price_c")

    def run(arg):
        out = subprocess.run([exe] + ([arg] if arg else []), capture_output=True, text=True, check=True).stdout
        keep = {}
        for ln in out.splitlines():
            for tag in ("C_ABI", "C_PIPE", "C_MORE"):
                if ln.startswith(tag):
                    keep[tag] = ln.split()[1:]
        return keep

    single = run(None)
    lists = ["0,0", "0,0,0,0,0"]
    n = torch.cuda.device_count()
    if n >= 2:
        lists.append(",".join(str(i) for i in range(min(n, 8))))
    for arg in lists:
        multi = run(arg)
        assert multi["C_ABI"] == single["C_ABI"], arg
        assert multi["C_MORE"] == single["C_MORE"], arg
        assert multi["C_PIPE"][0] == "0" and int(multi["C_PIPE"][1]) == len(arg.split(","))
        assert multi["C_PIPE"][2:] == single["C_PIPE"][2:], arg


def test_reference_hello_cu_on_all_shards(pkg):
    """The reference's call sequence (examples/hello_b200.cu = hello.cu over include/compat) opted into
    several shards with MCB200_DEVICES prints the same prices as on one."""
    exe = os.path.join(ROOT, "build", "hello_b200")
    if not os.path.exists(exe):
        pytest.skip("examples not built")

    def prices(env_value):
        env = dict(os.environ)
        if env_value:
            env["MCB200_DEVICES"] = env_value
        else:
            env.pop("MCB200_DEVICES", None)
        out = subprocess.run([exe], capture_output=True, text=True, check=True, env=env, timeout=900).stdout
        return [ln for ln in out.splitlines() if ln.startswith("Average GPU")]

    one = prices(None)
    assert len(one) >= 5
    assert prices("0,0,0") == one


def _silent_peer(conn):
    """A second process that exports a mailbox and then never submits a job."""
    import __graft_entry__ as entry
    pkg = entry.load_package()
    eng = pkg.Engine(0)
    conn.send((eng.peer_mailbox_create(), eng.peer_epoch()))
    conn.recv()          # the parent is done
    eng.close()


def test_a_missing_peer_poisons_the_result_instead_of_inventing_one(pkg, engine):
    """a peer that never delivers must not yield a plausible price.  Rank 0 of a two-rank process group
    (CUDA-IPC mailboxes) submits a job; rank 1 never does.  The final pass gives up at the %globaltimer deadline,
    writes NaN / n_paths = 0, collect returns MCB_ERR_TIMEOUT, and the engine prices normally again once it is back
    in a group of one.  (Safe on one GPU: the only kernel that waits is bounded and nothing else needs to run.)"""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    proc = ctx.Process(target=_silent_peer, args=(child,))
    proc.start()
    eng = pkg.Engine(0)
    try:
        peer_handle, peer_epoch = parent.recv()
        mine = eng.peer_mailbox_create()
        eng.peer_mailbox_connect(0, 2, [mine, peer_handle], max(eng.peer_epoch(), peer_epoch))
        eng.set_wait_timeout_ms(50)
        n = 100 * pkg.EUROPEAN_CHUNK + 7
        ticket = eng.european_submit(pkg.option(), n, 1234, pkg.CALL)
        with pytest.raises(pkg.McbError) as ei:
            eng.european_collect(ticket)
        assert ei.value.status == pkg.ERR_TIMEOUT
        assert eng.peer_timeouts() >= 1
        # a small job needs no peer (every rank prices it alone): it still comes out right in the broken group
        lone = eng.price_european(pkg.option(), 100_000, 1234, pkg.PUT)
        assert _bits(lone) == _bits(engine.price_european(pkg.option(), 100_000, 1234, pkg.PUT))
        # back to a group of one: same engine, same bits as ever
        eng.peer_mailbox_connect(0, 1, [mine], eng.peer_epoch())
        got = eng.price_european(pkg.option(), n, 1234, pkg.CALL)
        assert _bits(got) == _bits(engine.price_european(pkg.option(), n, 1234, pkg.CALL))
    finally:
        parent.send("done")
        proc.join(timeout=60)
        eng.close()
