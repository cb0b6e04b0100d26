"""GPU parity tests proper: the CUDA engine (through the C-ABI) against the CPU oracle, the golden
vectors, the cuRAND library stream and the closed forms.  Everything here needs a real B200.

Bars (see DESIGN.md "Parity"):
  * integer RNG words, path -> stream indexing, reduction trees: BIT-EXACT.
  * float normals / payoffs / prices: the engine computes in FP32 on the MUFU pipe, the oracle in
    double libm; tolerances are written next to each assert.
  * prices: within 3 standard errors of the closed form (north_star), SE from the engine itself.
"""
import ctypes as C
import os
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG1 = dict(S0=100.0, K=100.0, T=1.0, r=0.05, v=0.2)


# ------------------------------------------------------------------------------ integer stream
def test_philox_golden_vectors_on_device(engine, golden_philox):
    """Every stream vector of tests/golden/philox_vectors.json (cuRAND's header on the host, incl.
    the SURVEY 8(c) vectors) reproduced by the device kernel, bit for bit."""
    by_seed = {}
    for row in golden_philox["stream"]:
        by_seed.setdefault(int(row["seed"]), []).append(row)
    assert by_seed
    for seed, rs in by_seed.items():
        subs = [int(r["subsequence"]) for r in rs]
        blks = [int(r["block"]) for r in rs]
        got = engine.philox_blocks(seed, subs, blks)
        want = np.array([[int(w, 16) for w in r["out"]] for r in rs], dtype=np.uint32)
        assert (got == want).all()


def test_curand_normal_fixture(engine, golden_philox):
    """curand_normal() of the cuRAND Philox stream (host libm maths, committed fixture): the
    engine's normals agree to 4e-6 -- only the integer words are contractual."""
    for row in golden_philox["curand_normals"]:
        want = np.array(row["normals"], dtype=np.float64)
        got = engine.stream_normals(int(row["seed"]), int(row["subsequence"]), want.size)
        assert (np.abs(got - want) <= _normal_error_bound(want) + 2e-7).all()


def test_philox_matches_curand_library_on_device(engine, orc):
    """Live oracle of SURVEY 8(c): cuRAND's own curandStatePhilox4_32_10_t on the device,
    curand_init(seed, subsequence, 4*block) + curand4, word for word; and the CPU oracle."""
    rng = np.random.default_rng(7)
    n = 20000
    for seed in (1234, 1235, 0, 0xDEADBEEFCAFEF00D):
        subs = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
        subs[:8] = [0, 1, 2, 1023, 1024, (1 << 30) - 1, (1 << 32) - 1, (1 << 32) + 5]
        blks = rng.integers(0, 1 << 20, size=n, dtype=np.uint64)
        blks[:4] = [0, 1, 62, 63]
        mine = engine.philox_blocks(seed, subs, blks)
        theirs = engine.philox_blocks(seed, subs, blks, library=True)
        assert (mine == theirs).all()
        for i in range(0, n, 997):
            assert (mine[i] == orc.stream_block(seed, int(subs[i]), int(blks[i]))).all()


def _normal_error_bound(zo):
    """|dz| allowed between the engine's FP32/MUFU normals and the double oracle, per normal.
    With s the Box-Muller radius of the pair: MUFU.LG2 has absolute error <= 2^-22 on log2 u, so
    s^2 = -2 ln u is off by <= 3.3e-7 and s by <= 1.7e-7 / s (this is what dominates near u -> 1,
    where s -> 0); MUFU.SIN/COS on [-pi, pi) err <= 2^-21.4 and the FP32 angle carries 2^-24 * pi,
    both scaled by s; plus FP32 rounding of the product."""
    s = np.repeat(np.hypot(zo[0::2], zo[1::2]), 2)
    return 4e-7 + 8e-7 * s + 2e-7 / np.maximum(s, 1e-3)


def test_stream_normals_vs_oracle(engine, orc):
    """Engine normals (MUFU lg2/sqrt/sin/cos, FP32) vs oracle (libm, double) on the same words."""
    for seed, sub in ((1234, 0), (1234, 123456789), (1235, (1 << 40) + 17)):
        z = engine.stream_normals(seed, sub, 4096)
        zo = orc.stream_normals(seed, sub, 4096)
        assert np.isfinite(z).all()
        err = np.abs(z - zo)
        assert (err <= _normal_error_bound(zo)).all(), float((err / _normal_error_bound(zo)).max())
        assert np.quantile(err, 0.99) < 1.5e-6
    # offset start (n0 not a multiple of 4)
    z = engine.stream_normals(1234, 5, 1000, n0=3)
    zo = orc.stream_normals(1234, 5, 1004)
    assert (np.abs(z - zo[3:1003]) <= _normal_error_bound(zo)[3:1003]).all()


def test_normal_moments(engine):
    """E[z] = 0, E[z^2] = 1, E[z^3] = 0, E[z^4] = 3 over 2^22 normals of one stream (5 sigma)."""
    n = 1 << 22
    z = engine.stream_normals(99, 7, n).astype(np.float64)
    assert abs(z.mean()) < 5.0 / np.sqrt(n)
    assert abs((z ** 2).mean() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs((z ** 3).mean()) < 5.0 * np.sqrt(15.0 / n)
    assert abs((z ** 4).mean() - 3.0) < 5.0 * np.sqrt(96.0 / n)


# ------------------------------------------------------------------------------------ European
@pytest.mark.parametrize("kind", ["call", "put"])
def test_european_payoffs_vs_oracle(engine, orc, pkg, kind):
    """Per-path payoffs on the same (seed, path id) stream.  Tolerance 2e-4 absolute on payoffs of
    order 10-100: St = 2^(c0 + c1 z) in FP32 has relative error ~1e-6 (ex2.approx 2 ulp + the
    rounding of the exponent), i.e. ~1e-4 absolute at St ~ 100."""
    ot = pkg.PUT if kind == "put" else pkg.CALL
    for first, n in ((0, 40000), (16384 * 3 + 77, 20001), ((1 << 32) - 5000, 10000)):
        pay = engine.european_payoffs(pkg.option(**CFG1), first, n, 1234, ot)
        _, _, ref = orc.european(orc.option(**CFG1), first, n, 1234, ot, want_payoffs=True)
        assert np.abs(pay - ref).max() < 2e-4 * max(1.0, float(ref.max()) / 100.0), float(np.abs(pay - ref).max())
        assert ((pay == 0) == (ref == 0)).mean() > 0.9999


def test_european_reduction_tree_is_bit_exact(engine, orc, pkg):
    """Given the engine's own per-path payoffs, the chunk partials, the 64 double segments and the
    final sums must be BIT-identical to the oracle's restatement of the reduction tree."""
    n = 5 * pkg.EUROPEAN_CHUNK + 1234  # ragged tail
    opt = pkg.option(N_PATHS=n, **CFG1)
    pay = engine.european_payoffs(opt, 0, n, 1234, pkg.CALL)
    partials = engine.european_chunk_partials(opt, n, 1234, pkg.CALL)
    want = np.zeros_like(partials)
    for c in range(partials.shape[0]):
        lo = c * pkg.EUROPEAN_CHUNK
        cnt = min(pkg.EUROPEAN_CHUNK, n - lo)
        buf = np.zeros(pkg.EUROPEAN_CHUNK, np.float32)
        buf[:cnt] = pay[lo:lo + cnt]
        want[c] = orc.chunk_tree_f32(buf, cnt, pkg.EUROPEAN_PATHS_PER_SLOT)
    assert (partials.view(np.uint32) == want.view(np.uint32)).all()
    res = engine.price_european(opt, n, 1234, pkg.CALL)
    seg = engine.last_segments()
    seg_want = orc.segment_tree_f64(partials)
    assert (seg.view(np.uint64) == seg_want.view(np.uint64)).all()
    s, q = orc.final_tree_f64(seg)
    assert res.sum == s and res.sumsq == q and res.n_paths == n
    assert res.price == pytest.approx(orc.price_from_sum(s, n, opt.r, opt.T), rel=1e-15)
    assert res.std_error == pytest.approx(orc.std_error(s, q, n, opt.r, opt.T), rel=1e-12)


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 16383, 16384, 16385, 100_000, 1_000_000])
def test_european_sums_vs_oracle_ragged(engine, orc, pkg, n):
    """Edge sizes (single path, one slot row, chunk boundary +-1) and BASELINE configs[0] (1e6)."""
    for ot, oo in ((pkg.CALL, orc.CALL), (pkg.PUT, orc.PUT)):
        res = engine.price_european(pkg.option(N_PATHS=n, **CFG1), 0, 1234, ot)  # n_paths = 0 -> N_PATHS
        s, q = orc.european(orc.option(N_PATHS=n, **CFG1), 0, n, 1234, oo)
        assert res.n_paths == n
        # FP32 per-path error ~1e-6 relative to St, FP32 accumulation inside a 16384-path chunk
        assert abs(res.sum - s) <= 3e-5 * max(abs(s), 1.0) + 2e-4, (res.sum, s)
        assert abs(res.sumsq - q) <= 1e-4 * max(abs(q), 1.0) + 2e-2, (res.sumsq, q)


@pytest.mark.parametrize("log2n", [20, 26])
def test_european_within_3se_of_closed_form(engine, orc, pkg, log2n):
    n = 1 << log2n
    L = orc.lib()
    for ot, exact in ((pkg.CALL, L.orc_bs_call_exact(100, 100, 1, 0.05, 0.2)),
                      (pkg.PUT, L.orc_bs_put_exact(100, 100, 1, 0.05, 0.2))):
        res = engine.price_european(pkg.option(**CFG1), n, 1234, ot)
        assert abs(res.price - exact) < 3.0 * res.std_error, (res.price, exact, res.std_error)
        assert 0.5 < res.std_error * np.sqrt(n) / (14.0 if ot == pkg.CALL else 8.0) < 1.5


def test_european_full_size_2pow30(engine, orc, pkg):
    """BASELINE configs[1]: 2^30 paths, call and put, 3 SE of the closed form (SE ~ 4.5e-4), the
    reference's own float closed form, and put-call parity on common random numbers."""
    n = 1 << 30
    L = orc.lib()
    call = engine.price_european(pkg.option(**CFG1), n, 1234, pkg.CALL)
    put = engine.price_european(pkg.option(**CFG1), n, 1234, pkg.PUT)
    c_exact = L.orc_bs_call_exact(100, 100, 1, 0.05, 0.2)
    p_exact = L.orc_bs_put_exact(100, 100, 1, 0.05, 0.2)
    assert abs(call.price - c_exact) < 3.0 * call.std_error, (call.price, c_exact, call.std_error)
    assert abs(put.price - p_exact) < 3.0 * put.std_error, (put.price, p_exact, put.std_error)
    assert abs(call.price - L.orc_bs_call_reference(100, 100, 1, 0.05, 0.2)) < 3.0 * call.std_error + 1e-5
    # C - P = e^{-rT} E[St] - K e^{-rT}; Var[St] = S0^2 e^{2rT} (e^{sigma^2 T} - 1)
    fwd_se = 100.0 * np.sqrt(np.exp(0.04) - 1.0) / np.sqrt(n)
    assert abs((call.price - put.price) - (100.0 - 100.0 * np.exp(-0.05))) < 4.0 * fwd_se
    # same call twice -> same bits
    again = engine.price_european(pkg.option(**CFG1), n, 1234, pkg.CALL)
    assert again.sum == call.sum and again.sumsq == call.sumsq


def test_european_other_seeds_and_params(engine, orc, pkg):
    L = orc.lib()
    n = 1 << 22
    for seed in (1, 1235, 0x123456789ABCDEF):
        for S0, K, T, r, v in ((100, 120, 0.5, 0.01, 0.4), (50, 40, 2.0, 0.1, 0.15), (100, 100, 1.0, 0.1, 0.2)):
            opt = pkg.option(S0=S0, K=K, T=T, r=r, v=v)
            res = engine.price_european(opt, n, seed, pkg.CALL)
            exact = L.orc_bs_call_exact(S0, K, T, r, v)
            assert abs(res.price - exact) < 4.0 * res.std_error, (seed, S0, K, res.price, exact, res.std_error)


def _dev_array(torch, n, dtype):
    return torch.zeros(n, dtype=dtype, device="cuda:0")


def test_sharded_segments_bit_identical_for_any_world(engine, pkg):
    """One GPU stands in for 1/2/3/4/8 ranks: every rank fills only its own segments (+0.0
    elsewhere), the buffers are summed (what the NCCL allreduce does) and the fixed final tree
    runs.  sum / sumsq / price must be bit-identical for every world size."""
    import torch
    n = 37 * pkg.EUROPEAN_CHUNK + 999
    opt = pkg.option(N_PATHS=n, **CFG1)
    ref = engine.price_european(opt, n, 1234, pkg.CALL)
    res_dev = torch.zeros(5, dtype=torch.float64, device="cuda:0")  # one mcb_result (40 bytes)
    for world in (1, 2, 3, 4, 8):
        total = _dev_array(torch, 2 * pkg.SEGMENTS, torch.float64)
        for rank in range(world):
            seg = _dev_array(torch, 2 * pkg.SEGMENTS, torch.float64)
            engine.european_segments_async(opt, n, 1234, pkg.CALL, rank, world, seg.data_ptr())
            engine.synchronize()
            total += seg
        torch.cuda.synchronize()
        engine.combine_segments_async(total.data_ptr(), 1, n, opt.r, opt.T, res_dev.data_ptr())
        engine.synchronize()
        host = res_dev.cpu().numpy()
        assert host[2] == ref.sum and host[3] == ref.sumsq and host[0] == ref.price and host[1] == ref.std_error


# -------------------------------------------------------------------------------------- sweep
def test_sweep_is_bit_identical_to_separate_calls(engine, orc, pkg):
    n = 3 * pkg.EUROPEAN_CHUNK + 5
    strikes = np.linspace(60, 140, 6, dtype=np.float32)
    vols = np.linspace(0.05, 0.8, 6, dtype=np.float32)
    K, V = np.meshgrid(strikes, vols, indexing="ij")
    for ot in (pkg.CALL, pkg.PUT):
        out = engine.price_sweep(pkg.option(**CFG1), K.ravel(), V.ravel(), n, 1234, ot)
        for i, (k, v) in enumerate(zip(K.ravel(), V.ravel())):
            one = engine.price_european(pkg.option(S0=100.0, T=1.0, r=0.05, K=float(k), v=float(v)), n, 1234, ot)
            assert out[i].sum == one.sum and out[i].sumsq == one.sumsq and out[i].price == one.price
    s, q = orc.sweep(orc.option(**CFG1), K.ravel(), V.ravel(), 0, n, 1234, orc.CALL)
    out = engine.price_sweep(pkg.option(**CFG1), K.ravel(), V.ravel(), n, 1234, pkg.CALL)
    for i in range(K.size):
        assert abs(out[i].sum - s[i]) <= 5e-5 * max(abs(s[i]), 1.0) + 1e-3


def test_sweep_closed_form(engine, orc, pkg):
    L = orc.lib()
    n = 1 << 22
    strikes = np.linspace(60, 140, 8, dtype=np.float32)
    vols = np.linspace(0.05, 0.8, 8, dtype=np.float32)
    K, V = np.meshgrid(strikes, vols, indexing="ij")
    out = engine.price_sweep(pkg.option(**CFG1), K.ravel(), V.ravel(), n, 1234, pkg.CALL)
    z = []
    for i, (k, v) in enumerate(zip(K.ravel(), V.ravel())):
        exact = L.orc_bs_call_exact(100.0, float(k), 1.0, 0.05, float(v))
        d2 = (np.log(100.0 / k) + (0.05 - 0.5 * v * v)) / v
        p_itm = 0.5 * math.erfc(-d2 / np.sqrt(2.0))
        if n * p_itm < 1000:   # (almost) no path ends in the money: price ~ 0 +- 0, nothing to z-test
            assert out[i].price <= exact + 1e-3
            continue
        z.append((out[i].price - exact) / out[i].std_error)
    z = np.array(z)
    # common random numbers: the z-scores are correlated, so bound each one (4 SE) rather than chi^2
    assert z.size >= 50 and np.abs(z).max() < 4.0, z


# ------------------------------------------------------------------------------------- bullet
BUL = dict(S0=100.0, K=100.0, T=1.0, r=0.05, v=0.2, B=120.0, P1=10, P2=50)


def test_bullet_payoffs_vs_oracle(engine, orc, pkg):
    """Multi-step walk + barrier count.  The count is an integer function of FP32 comparisons
    (log2 S < log2 B); a path whose log-price passes within FP32 rounding of the barrier can count
    one step differently from the double oracle, which flips its payoff gate only when the count
    sits exactly on P1 / P2.  So: >= 99.9 % of payoffs agree to 1e-3 relative, none is non-finite."""
    for n_steps, n in ((100, 6000), (7, 5000), (1, 3000), (252, 2000)):
        kw = dict(N_STEPS=n_steps, N_PATHS=n, **BUL)
        if n_steps < 60:
            kw.update(P1=1, P2=n_steps)
        pay = engine.bullet_payoffs(pkg.option(**kw), 0, n, 1234)
        _, _, ref = orc.bullet(orc.option(**kw), 0, n, 1234, want_payoffs=True)
        assert np.isfinite(pay).all()
        close = np.abs(pay - ref) <= 1e-3 * np.maximum(ref, 1.0)
        assert close.mean() >= 0.999, (n_steps, float(close.mean()))


def test_bullet_restart_state(engine, orc, pkg):
    kw = dict(N_STEPS=100, N_PATHS=4000, **BUL)
    pay = engine.bullet_payoffs(pkg.option(**kw), 100, 4000, 1234, Ik=7, Sk=110.0, Tk=40)
    _, _, ref = orc.bullet(orc.option(**kw), 100, 4000, 1234, Ik=7, Sk=110.0, Tk=40, want_payoffs=True)
    close = np.abs(pay - ref) <= 1e-3 * np.maximum(ref, 1.0)
    assert close.mean() >= 0.999
    # Tk == N_STEPS: no steps left, payoff is the gated intrinsic value of Sk
    pay = engine.bullet_payoffs(pkg.option(**kw), 0, 10, 1234, Ik=12, Sk=130.0, Tk=100)
    assert np.allclose(pay, 30.0, rtol=1e-5)
    pay = engine.bullet_payoffs(pkg.option(**kw), 0, 10, 1234, Ik=3, Sk=130.0, Tk=100)
    assert (pay == 0).all()


def test_bullet_price_and_tree(engine, orc, pkg):
    n = 50_000
    kw = dict(N_STEPS=100, N_PATHS=n, **BUL)
    opt = pkg.option(**kw)
    res = engine.price_bullet(opt, n, 1234)
    s, q = orc.bullet(orc.option(**kw), 0, n, 1234)
    se_sum = np.sqrt(max(q - s * s / n, 0.0))
    assert abs(res.sum - s) <= 0.05 * se_sum + 1e-4 * s, (res.sum, s, se_sum)
    # the engine's own payoffs through the oracle's restatement of the tree -> bit-identical sums
    pay = engine.bullet_payoffs(opt, 0, n, 1234)
    nchunks = (n + pkg.BULLET_CHUNK - 1) // pkg.BULLET_CHUNK
    partials = np.zeros((nchunks, 2), np.float32)
    for c in range(nchunks):
        lo = c * pkg.BULLET_CHUNK
        cnt = min(pkg.BULLET_CHUNK, n - lo)
        buf = np.zeros(pkg.BULLET_CHUNK, np.float32)
        buf[:cnt] = pay[lo:lo + cnt]
        partials[c] = orc.chunk_tree_f32(buf, cnt, pkg.BULLET_PATHS_PER_SLOT)
    s2, q2 = orc.final_tree_f64(orc.segment_tree_f64(partials))
    assert res.sum == s2 and res.sumsq == q2


def test_bullet_without_barrier_is_european(engine, orc, pkg):
    """B = 0, P1 = 0: the gate is always open, so the 64-step bullet price is a European call."""
    n = 1 << 21
    opt = pkg.option(S0=100.0, K=100.0, T=1.0, r=0.05, v=0.2, B=0.0, P1=0, P2=1000, N_STEPS=64, N_PATHS=n)
    res = engine.price_bullet(opt, n, 1234)
    exact = orc.lib().orc_bs_call_exact(100, 100, 1, 0.05, 0.2)
    assert abs(res.price - exact) < 3.0 * res.std_error, (res.price, exact, res.std_error)


# -------------------------------------------------------------------------------- trajectories
@pytest.mark.parametrize("n_steps,n_paths,first", [(252, 300, 0), (100, 129, 1000), (7, 33, 5), (1, 64, 0),
                                                   (33, 31, (1 << 32) - 16), (64, 1, 9), (32, 40, 0), (150, 20, 0),
                                                   (192, 9, 7), (300, 9, 0), (600, 5, 3), (1023, 3, 1), (5000, 2, 0),
                                                   (1100, 7, 2), (2048, 3, 0)])
def test_trajectories_vs_oracle(engine, orc, pkg, n_steps, n_paths, first):
    """Path-major prices[p][i] = S(t_{i+1}) and barrier counts.  FP32 log2-space accumulation over
    n_steps steps: relative error <= ~n_steps * 2^-24 * |log2 S| ~ 1e-4 at 252 steps -> rtol 3e-4.
    Counts are integers: a mismatch is only tolerated where the oracle's price is within 3e-4
    relative of the barrier at some step (FP32 rounding of the comparison)."""
    kw = dict(N_STEPS=n_steps, N_PATHS=n_paths, **BUL)
    prices, counts = engine.simulate_trajectories(pkg.option(**kw), first, n_paths, 1234, want_counts=True)
    rp, rc = orc.trajectories(orc.option(**kw), first, n_paths, 1234)
    assert prices.shape == (n_paths, n_steps)
    assert np.allclose(prices, rp, rtol=3e-4, atol=0), float(np.abs(prices / rp - 1).max())
    bad = np.nonzero((counts != rc).any(axis=1))[0]
    for p in bad:
        assert (np.abs(rp[p] / 120.0 - 1.0) < 3e-4).any(), p
    only = engine.simulate_trajectories(pkg.option(**kw), first, n_paths, 1234)
    assert (only.view(np.uint32) == prices.view(np.uint32)).all()  # same bits with and without counts


def test_trajectories_are_independent_of_the_launch_slab(engine, pkg):
    """Path p's row is a pure function of (seed, p): slabs cut anywhere give the same bits
    (this is what makes trajectory mode shard across GPUs with no collective)."""
    kw = dict(N_STEPS=252, N_PATHS=1000, **BUL)
    full = engine.simulate_trajectories(pkg.option(**kw), 0, 1000, 1234)
    for lo, hi in ((0, 500), (500, 1000), (37, 38), (333, 777)):
        part = engine.simulate_trajectories(pkg.option(**kw), lo, hi - lo, 1234)
        assert (part.view(np.uint32) == full[lo:hi].view(np.uint32)).all()


def test_trajectory_terminal_matches_bullet_walk(engine, pkg):
    """The last stored price of path p and the bullet kernel's terminal payoff come from the same
    normals; the trajectory kernel sums the log2 increments by in-lane prefix + warp scan, the
    bullet kernel serially, so they agree to FP32 rounding of a 100-term sum (~5e-6 relative,
    i.e. < 2e-3 absolute on prices ~100-200)."""
    kw = dict(S0=100.0, K=100.0, T=1.0, r=0.05, v=0.2, B=0.0, P1=0, P2=1000, N_STEPS=100, N_PATHS=2048)
    prices = engine.simulate_trajectories(pkg.option(**kw), 0, 2048, 1234)
    pay = engine.bullet_payoffs(pkg.option(**kw), 0, 2048, 1234)
    assert np.allclose(pay, np.maximum(prices[:, -1] - 100.0, 0.0), rtol=0, atol=2e-3)


def test_trajectories_device_buffer_full_size(engine, pkg):
    """BASELINE configs[2]: 2^20 paths x 252 steps into a device buffer; size-independent checks:
    every value finite and positive, E[S_T] = S0 e^{rT} within 4 SE, sampled rows equal a small
    re-run bit for bit."""
    import torch
    n, steps = 1 << 20, 252
    opt = pkg.option(N_STEPS=steps, N_PATHS=n, **BUL)
    buf = torch.empty(n * steps, dtype=torch.float32, device="cuda:0")
    engine.trajectories_async(opt, 0, n, 1234, buf.data_ptr())
    engine.synchronize()
    m = buf.view(n, steps)
    assert bool(torch.isfinite(m).all()) and bool((m > 0).all())
    st = m[:, -1].double()
    se = float(st.std()) / np.sqrt(n)
    assert abs(float(st.mean()) - 100.0 * np.exp(0.05)) < 4.0 * se
    for lo in (0, 77777, n - 64):
        part = engine.simulate_trajectories(opt, lo, 64, 1234)
        assert (part.view(np.uint32) == m[lo:lo + 64].cpu().numpy().view(np.uint32)).all()


# ----------------------------------------------------------------------------------- nested MC
def test_nested_vs_oracle_small(engine, orc, pkg):
    """F[p,k] against the oracle on the same outer/inner streams.  Each F is a mean of N_inner
    gated payoffs; FP32 count flips change single inner paths, so compare with a tolerance of
    0.5 % of the point's scale + 2 gated payoffs' worth."""
    kw = dict(N_STEPS=20, N_PATHS=3, N_PATHS_INNER=300, **{**BUL, "P1": 2, "P2": 12})
    for mode_e, mode_o in ((pkg.DISCOUNT_COMPAT, orc.DISCOUNT_COMPAT), (pkg.DISCOUNT_CORRECT, orc.DISCOUNT_CORRECT)):
        F, P, Cn, mean = engine.nested_monte_carlo(pkg.option(**kw), 5, 3, 1234, 1235, mode_e)
        Fo, Po, Co = orc.nmc(orc.option(**kw), 5, 3, 1234, 1235, mode_o)
        assert np.allclose(P, Po, rtol=1e-4)
        assert (Cn == Co).mean() > 0.95
        assert np.abs(F - Fo).max() <= 5e-3 * max(1.0, float(Fo.max())) + 2.0 * 50.0 / 300, float(np.abs(F - Fo).max())
        assert mean == pytest.approx(float(F.astype(np.float64).sum()) / (F.size + 1), rel=1e-12)


def test_nested_last_step_is_the_gated_intrinsic_value(engine, pkg):
    kw = dict(N_STEPS=16, N_PATHS=64, N_PATHS_INNER=512, **{**BUL, "P1": 0, "P2": 16})
    F, P, Cn, _ = engine.nested_monte_carlo(pkg.option(**kw), 0, 64, 1234, 1235, pkg.DISCOUNT_CORRECT)
    want = np.maximum(P[:, -1] - 100.0, 0.0)  # tau = 0 at the last step -> discount 1
    assert np.allclose(F[:, -1], want, rtol=1e-5, atol=1e-4)


def test_nested_without_barrier_matches_black_scholes(engine, orc, pkg):
    """B = 0, P1 = 0 in CORRECT discount mode: F[p,k] -> C(S[p,k], K, T - t_{k+1}) (SURVEY 8(c) NMC
    check ii).  4096 inner paths: per-point SE ~ sd/64; require |z| < 4.5 on every point and a
    mean z near 0."""
    L = orc.lib()
    steps, n_in = 10, 4096
    kw = dict(S0=100.0, K=100.0, T=1.0, r=0.05, v=0.2, B=0.0, P1=0, P2=1000, N_STEPS=steps, N_PATHS=8,
              N_PATHS_INNER=n_in)
    F, P, _, _ = engine.nested_monte_carlo(pkg.option(**kw), 0, 8, 1234, 1235, pkg.DISCOUNT_CORRECT)
    zs = []
    for p in range(8):
        for k in range(steps - 1):
            tau = 1.0 - (k + 1) * 0.1
            exact = L.orc_bs_call_exact(float(P[p, k]), 100.0, tau, 0.05, 0.2)
            sd = 1.6 * max(exact, 0.5) + 2.0  # generous payoff sd proxy
            zs.append((F[p, k] - exact) / (sd / np.sqrt(n_in)))
    zs = np.array(zs)
    assert np.abs(zs).max() < 4.5 and abs(zs.mean()) < 0.6, (np.abs(zs).max(), zs.mean())


def test_nested_outer_paths_match_trajectory_mode(engine, pkg):
    kw = dict(N_STEPS=40, N_PATHS=16, N_PATHS_INNER=64, **BUL)
    _, P, Cn, _ = engine.nested_monte_carlo(pkg.option(**kw), 3, 16, 1234, 1235, pkg.DISCOUNT_COMPAT)
    tp, tc = engine.simulate_trajectories(pkg.option(**kw), 3, 16, 1234, want_counts=True)
    assert (P.view(np.uint32) == tp.view(np.uint32)).all() and (Cn == tc).all()


# --------------------------------------------------------------- reduce / pre-generated normals
def test_reduce_sum_bit_exact(engine, orc):
    rng = np.random.default_rng(3)
    for n in (0, 1, 255, 256, 1024, 100_003):
        x = rng.standard_normal(n).astype(np.float32)
        assert engine.reduce_sum(x) == orc.reduce_sum_f32(x)


def test_price_from_pregenerated_normals(engine, orc, pkg):
    rng = np.random.default_rng(5)
    z = rng.standard_normal((5000, 12)).astype(np.float32)
    kw = dict(N_STEPS=12, **CFG1)
    pay = engine.price_from_normals(pkg.option(**kw), z)
    ref = orc.pregen_european(orc.option(**kw), z)
    assert np.abs(pay - ref).max() < 5e-4


# ------------------------------------------------------------------- boundary: errors, wrappers
def test_invalid_arguments_are_status_codes_not_exits(engine, pkg):
    lib = pkg.load_library()
    out = pkg.Result()
    bad = pkg.option(S0=-1.0)
    assert lib.mcb_price_european(engine._h, C.byref(bad), 10, 1, pkg.CALL, C.byref(out)) == pkg.ERR_INVALID
    assert b"S0" in lib.mcb_last_error()
    ok = pkg.option()
    assert lib.mcb_price_european(engine._h, C.byref(ok), 10, 1, 7, C.byref(out)) == pkg.ERR_INVALID
    assert lib.mcb_price_european(engine._h, C.byref(pkg.option(N_PATHS=0)), 0, 1, 0, C.byref(out)) == pkg.ERR_INVALID
    assert lib.mcb_price_european(None, C.byref(ok), 10, 1, 0, C.byref(out)) == pkg.ERR_INVALID
    assert lib.mcb_price_bullet(engine._h, C.byref(pkg.option(N_STEPS=0)), 10, 1, 0, 0.0, 0, C.byref(out)) == \
        pkg.ERR_INVALID
    with pytest.raises(pkg.McbError):
        engine.price_bullet(pkg.option(N_STEPS=10), 10, 1, Tk=11)
    # the engine is still usable after errors
    assert engine.price_european(ok, 1000, 1234, pkg.CALL).n_paths == 1000


def test_reference_wrappers_call_surface(pkg, orc, capsys):
    """inc/wrappers.cuh names, float return, stdout side effect, threadsPerBlock ignored."""
    opt = pkg.option(r=0.1, N_PATHS=100000, N_PATHS_INNER=64, N_STEPS=100)  # hello.cu:5-17 parameters
    a = pkg.wrapper_gpu_option_vanilla(opt, 1024)
    b = pkg.wrapper_gpu_option_vanilla(opt, 128, quiet=True)
    assert a == b
    assert "Average GPU" in capsys.readouterr().out
    bs = orc.lib().orc_bs_call_reference(100, 100, 1, 0.1, 0.2)
    assert abs(a - bs) < 0.2  # 1e5 paths: SE ~ 0.05
    c = pkg.wrapper_gpu_bullet_option(opt, 1024, quiet=True)
    d = pkg.wrapper_gpu_bullet_option_atomic(opt, 1024, quiet=True)
    assert c == d and 0.0 < c < a
    small = pkg.option(r=0.1, N_PATHS=32, N_PATHS_INNER=64, N_STEPS=20)
    e = pkg.wrapper_gpu_bullet_option_nmc_one_point_one_block(small, 1024, 5000, quiet=True)
    f = pkg.wrapper_gpu_bullet_option_nmc_one_kernel(small, 1024, 5000, quiet=True)
    g = pkg.wrapper_gpu_bullet_option_nmc_optimal(small, 1024, 5000, quiet=True)
    assert e == f == g and e >= 0.0


def test_device_info(engine):
    info = engine.device_info()
    assert info.cc_major == 10 and info.sm_count >= 100
    assert b"B200" in info.name or b"NVIDIA" in info.name


# ------------------------------------------------------------- out-of-bounds canaries (no sanitizer on this pool)
@pytest.mark.parametrize("n_steps,n_paths", [(252, 61), (252, 24), (256, 7), (100, 33), (128, 5), (7, 5), (300, 9),
                                             (1, 3), (37, 1), (1024, 3),
                                             # rows too long to stage whole: pass-by-pass staging (trajectory_long_kernel);
                                             # an odd row count and a ragged last pass (1100 = 4 x 256 + 76)
                                             (2048, 5), (1100, 3), (1536, 2), (4100, 1)])
def test_trajectory_kernels_stay_inside_their_buffers(engine, pkg, n_steps, n_paths):
    """compute-sanitizer is closed on this pool, so bounds are checked with guard bands: the kernels
    (TMA slab kernel for single-pass aligned rows, general kernel otherwise, with and without
    counts) must write every element of [guard, guard + n) and nothing else."""
    import torch
    guard = 4096
    n = n_steps * n_paths
    opt = pkg.option(N_STEPS=n_steps, N_PATHS=n_paths, B=120.0)
    for with_counts in (False, True):
        prices = torch.full((n + 2 * guard,), float("nan"), dtype=torch.float32, device="cuda:0")
        counts = torch.full((n + 2 * guard,), -7, dtype=torch.int32, device="cuda:0") if with_counts else None
        engine.trajectories_async(opt, 11, n_paths, 1234, prices[guard:].data_ptr(),
                                  counts[guard:].data_ptr() if with_counts else None)
        engine.synchronize()
        assert bool(torch.isnan(prices[:guard]).all()) and bool(torch.isnan(prices[guard + n:]).all())
        assert bool(torch.isfinite(prices[guard:guard + n]).all())
        if with_counts:
            assert bool((counts[:guard] == -7).all()) and bool((counts[guard + n:] == -7).all())
            assert bool((counts[guard:guard + n] >= 0).all())
        host = engine.simulate_trajectories(opt, 11, n_paths, 1234)
        assert (prices[guard:guard + n].cpu().numpy().view(np.uint32) == host.ravel().view(np.uint32)).all()


def test_nested_and_segment_outputs_stay_inside_their_buffers(engine, pkg):
    import torch
    guard = 1024
    opt = pkg.option(N_STEPS=13, N_PATHS=5, N_PATHS_INNER=300, B=120.0, P1=1, P2=10)
    n = 13 * 5
    F = torch.full((n + 2 * guard,), float("nan"), dtype=torch.float32, device="cuda:0")
    engine.nested_async(opt, 2, 5, 1234, 1235, pkg.DISCOUNT_CORRECT, F[guard:].data_ptr())
    engine.synchronize()
    assert bool(torch.isnan(F[:guard]).all()) and bool(torch.isnan(F[guard + n:]).all())
    assert bool(torch.isfinite(F[guard:guard + n]).all())
    seg = torch.full((2 * pkg.SEGMENTS + 2 * guard,), float("nan"), dtype=torch.float64, device="cuda:0")
    engine.european_segments_async(pkg.option(), 100001, 1234, pkg.CALL, 0, 1, seg[guard:].data_ptr())
    engine.synchronize()
    assert bool(torch.isnan(seg[:guard]).all()) and bool(torch.isnan(seg[guard + 2 * pkg.SEGMENTS:]).all())
    assert bool(torch.isfinite(seg[guard:guard + 2 * pkg.SEGMENTS]).all())


# ------------------------------------------------------------------ BASELINE configs[3] / [4] at full size
def test_sweep_full_size_config5(engine, orc, pkg):
    """BASELINE configs[4] on one GPU: 1024 parameter sets (32 strikes x 32 vols) x 2^26 paths with
    common random numbers.  Every set with enough in-the-money mass within 4 SE of the closed form;
    prices monotone in the strike for every vol; a spot-checked set bit-identical to a separate
    European call."""
    n = 1 << 26
    strikes = np.linspace(60, 140, 32, dtype=np.float32)
    vols = np.linspace(0.05, 0.8, 32, dtype=np.float32)
    K, V = np.meshgrid(strikes, vols, indexing="ij")
    out = engine.price_sweep(pkg.option(**CFG1), K.ravel(), V.ravel(), n, 1234, pkg.CALL)
    L = orc.lib()
    price = np.array([r.price for r in out]).reshape(32, 32)
    checked = 0
    for i, (k, v) in enumerate(zip(K.ravel(), V.ravel())):
        d2 = (np.log(100.0 / k) + (0.05 - 0.5 * v * v)) / v
        if n * 0.5 * math.erfc(-d2 / np.sqrt(2.0)) < 1e5:
            continue
        exact = L.orc_bs_call_exact(100.0, float(k), 1.0, 0.05, float(v))
        assert abs(out[i].price - exact) < 4.0 * out[i].std_error, (k, v, out[i].price, exact, out[i].std_error)
        checked += 1
    assert checked > 900
    assert (np.diff(price, axis=0) <= 1e-9).all()            # call price decreases with the strike (same draws)
    one = engine.price_european(pkg.option(S0=100.0, T=1.0, r=0.05, K=float(K[17, 9]), v=float(V[17, 9])), n, 1234)
    assert out[17 * 32 + 9].sum == one.sum and out[17 * 32 + 9].sumsq == one.sumsq


def test_nested_full_size_config4(engine, orc, pkg):
    """BASELINE configs[3]: 4096 outer x 4096 inner x 100 steps (8.3e10 inner path-steps) on device
    buffers.  Size-independent checks: F finite and >= 0; the last step is the gated intrinsic value;
    rows re-run as a small job come back bit-identical (CTA p depends only on p); tower property:
    the mean over outer paths of F[:, k] (compat discount = e^{-rT} for every k) is the time-0 bullet
    price for every k, within 4 SE of its cross-sectional spread."""
    import torch
    n_out, n_in, steps = 4096, 4096, 100
    opt = pkg.option(N_STEPS=steps, N_PATHS=n_out, N_PATHS_INNER=n_in, B=120.0, P1=10, P2=50, **CFG1)
    F = torch.empty(n_out * steps, dtype=torch.float32, device="cuda:0")
    P = torch.empty(n_out * steps, dtype=torch.float32, device="cuda:0")
    Cn = torch.empty(n_out * steps, dtype=torch.int32, device="cuda:0")
    engine.nested_async(opt, 0, n_out, 1234, 1235, pkg.DISCOUNT_COMPAT, F.data_ptr(), P.data_ptr(), Cn.data_ptr())
    engine.synchronize()
    Fh = F.view(n_out, steps).cpu().numpy().astype(np.float64)
    Ph = P.view(n_out, steps).cpu().numpy()
    Ch = Cn.view(n_out, steps).cpu().numpy()
    assert np.isfinite(Fh).all() and (Fh >= 0).all()
    gate = (Ch[:, -1] >= 10) & (Ch[:, -1] <= 50)
    want_last = np.where(gate, np.maximum(Ph[:, -1] - 100.0, 0.0), 0.0) * np.exp(-0.05)
    assert np.allclose(Fh[:, -1], want_last, rtol=1e-5, atol=1e-4)
    small, _, _, _ = engine.nested_monte_carlo(pkg.option(N_STEPS=steps, N_PATHS=2, N_PATHS_INNER=n_in, B=120.0,
                                                          P1=10, P2=50, **CFG1), 1000, 2, 1234, 1235,
                                               pkg.DISCOUNT_COMPAT)
    assert (small.view(np.uint32) == F.view(n_out, steps)[1000:1002].cpu().numpy().view(np.uint32)).all()
    ref = engine.price_bullet(pkg.option(N_STEPS=steps, N_PATHS=1 << 22, B=120.0, P1=10, P2=50, **CFG1), 1 << 22, 99)
    for k in (0, 10, 50, 90, 99):
        col = Fh[:, k]
        se = col.std() / np.sqrt(n_out)
        assert abs(col.mean() - ref.price) < 4.0 * np.hypot(se, ref.std_error), (k, col.mean(), ref.price, se)


def test_boxmuller_maps_exhaustively(engine):
    """Every one of the 2^32 possible words through the MUFU radius / sin / cos maps against double
    precision: no NaN, no negative radius (log(0) cannot happen: u in (0, 1]), and the worst-case
    absolute errors that the per-normal tolerances above are derived from.
      radius: MUFU.LG2 absolute error 2^-22 on log2 u blows up as 1.7e-7 / s near u -> 1, where
              s -> 0; the largest error is therefore at the smallest non-zero radius (~3.5e-4);
      sin/cos: MUFU on [-pi, pi) plus the FP32 angle, <= ~1e-6."""
    err_r, bad_r = engine.boxmuller_scan(0)
    err_s, bad_s = engine.boxmuller_scan(1)
    err_c, bad_c = engine.boxmuller_scan(2)
    assert bad_r == 0 and bad_s == 0 and bad_c == 0
    assert err_r < 4e-4, err_r
    assert err_s < 1.5e-6 and err_c < 1.5e-6, (err_s, err_c)
    # away from u -> 1 (x < 2^32 - 2^24, i.e. s > ~0.09) the radius is good to ~2e-6
    err_bulk, _ = engine.boxmuller_scan(0, 0, (1 << 32) - (1 << 24))
    assert err_bulk < 3e-6, err_bulk


def test_sweep_with_more_sets_than_a_grid_dimension(engine, pkg):
    """70 000 parameter sets on a tiny path count: the sets are priced in groups (gridDim.y <= 65535)."""
    n_sets = 70_000
    k = np.linspace(50, 150, n_sets, dtype=np.float32)
    v = np.full(n_sets, 0.2, dtype=np.float32)
    out = engine.price_sweep(pkg.option(**CFG1), k, v, 5000, 1234, pkg.CALL)
    assert len(out) == n_sets
    for i in (0, 1, 65534, 65535, 65536, n_sets - 1):
        one = engine.price_european(pkg.option(S0=100.0, T=1.0, r=0.05, K=float(k[i]), v=0.2), 5000, 1234, pkg.CALL)
        assert out[i].sum == one.sum and out[i].sumsq == one.sumsq


def test_gpu_prices_agree_with_the_reference_cpu_path(engine, orc, pkg):
    """north_star: "prices agree with the reference's CPU path ... within 3 standard errors".  The
    UNMODIFIED reference CPU pricers (oracle/_ref, inc/tool.cuh:104-173) are unseeded, so the
    comparison is statistical: 8 reference runs of 2^20 paths against the engine's 2^26-path price
    (whose own SE is negligible next to the reference's).  The reference seeds from random_device, so
    the bound used is 4 SE (a 3 SE bound would fail one honest run in ~200)."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref is built only where /root/reference exists")
    ref = orc.ref_cpu()
    n_ref, runs = 1 << 20, 8
    # vanilla (hello.cu parameters, r = 0.1)
    o = orc.option(r=0.1, N_PATHS=n_ref)
    prices = np.array([ref.ref_vanilla_cpu(C.byref(o)) for _ in range(runs)], dtype=np.float64)
    gpu = engine.price_european(pkg.option(r=0.1), 1 << 26, 1234, pkg.CALL)
    se_ref = gpu.std_error * np.sqrt((1 << 26) / (n_ref * runs))     # SE of the pooled reference mean
    assert abs(prices.mean() - gpu.price) < 4.0 * np.hypot(se_ref, gpu.std_error) + 2e-3, (prices, gpu)
    # bullet (hello.cu parameters)
    ob = orc.option(r=0.1, B=120.0, P1=10, P2=50, N_STEPS=100, N_PATHS=1 << 17)
    bp = np.array([ref.ref_bullet_cpu(C.byref(ob)) for _ in range(runs)], dtype=np.float64)
    gb = engine.price_bullet(pkg.option(r=0.1, B=120.0, P1=10, P2=50, N_STEPS=100), 1 << 22, 1234)
    se_b = gb.std_error * np.sqrt((1 << 22) / ((1 << 17) * runs))
    assert abs(bp.mean() - gb.price) < 4.0 * np.hypot(se_b, gb.std_error) + 2e-3, (bp, gb)


def test_engine_vs_the_reference_gpu_wrappers(engine, orc, pkg):
    """SURVEY 8(c) live oracle: the reference's OWN GPU wrappers (unmodified inc/*.cuh compiled for
    sm_100 into oracle/_ref/ref_gpu in the build container; XORWOW + float atomics) run on this B200
    next to the engine.  Different generators, so the comparison is statistical: 2^20 paths each,
    4 SE of the difference.  An absurd reference price means its never-zeroed accumulator
    (inc/wrappers.cuh:43-47) picked up stale memory: reported as xfail, not as a mismatch."""
    import subprocess
    exe = os.path.join(os.path.dirname(orc.__file__), "_ref", "ref_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gpu is built only where /root/reference exists")
    n = 1 << 20
    out = subprocess.run([exe, str(n), "0.1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    tag = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("REFGPU")][0]
    ref_vanilla, ref_bullet = float(tag[2]), float(tag[3])
    mine_v = engine.price_european(pkg.option(r=0.1), n, 1234, pkg.CALL)
    mine_b = engine.price_bullet(pkg.option(r=0.1, B=120.0, P1=10, P2=50, N_STEPS=100), n, 1234)
    if not (0.0 < ref_vanilla < 1e3 and 0.0 <= ref_bullet < 1e3):
        pytest.xfail(f"reference accumulators not zeroed: {ref_vanilla}, {ref_bullet}")
    # both estimators have the engine's SE at this N (same payoff distribution): difference ~ sqrt(2) SE
    assert abs(ref_vanilla - mine_v.price) < 4.0 * np.sqrt(2.0) * mine_v.std_error + 1e-3, (ref_vanilla, mine_v)
    assert abs(ref_bullet - mine_b.price) < 4.0 * np.sqrt(2.0) * mine_b.std_error + 1e-3, (ref_bullet, mine_b)


def test_trajectories_to_host_in_slabs(engine, pkg):
    """A host-destination request larger than the 128 MB workspace slab (2^18 rows x 252 steps = 264 MB,
    + counts) is produced slab by slab; rows must not depend on where the slab boundaries fall."""
    n, steps = 1 << 18, 252
    opt = pkg.option(N_STEPS=steps, N_PATHS=n, B=120.0)
    prices, counts = engine.simulate_trajectories(opt, 5, n, 1234, want_counts=True)
    assert np.isfinite(prices).all() and (counts >= 0).all() and (np.diff(counts, axis=1) >= 0).all()
    for lo in (0, 133150, 133160, n - 10):      # 133 152 rows per slab: straddle the first boundary
        p2, c2 = engine.simulate_trajectories(opt, 5 + lo, 10, 1234, want_counts=True)
        assert (p2.view(np.uint32) == prices[lo:lo + 10].view(np.uint32)).all() and (c2 == counts[lo:lo + 10]).all()


def test_european_deep_bias_check_2pow34(engine, orc, pkg):
    """Fast-math bias (SURVEY section 7 "hard parts"): 2^34 paths, SE = 1.1e-4 on the call -- any
    systematic error of the MUFU / FP32 pipeline above ~3e-5 relative would show as |z| > 3.  Measured
    offline at 2^36 paths (SE 5.6e-5) for three seeds: z = -0.35, -0.92, -0.82 (call), -0.23, -0.82,
    +0.01 (put), i.e. a bias, if any, below 5e-6 relative (the FP32 rounding of the folded exponent
    constants c0, c1 is of that order)."""
    n = 1 << 34
    L = orc.lib()
    for ot, exact in ((pkg.CALL, L.orc_bs_call_exact(100, 100, 1, 0.05, 0.2)),
                      (pkg.PUT, L.orc_bs_put_exact(100, 100, 1, 0.05, 0.2))):
        res = engine.price_european(pkg.option(**CFG1), n, 20261018, ot)
        assert res.n_paths == n
        assert abs(res.price - exact) < 3.0 * res.std_error, (res.price, exact, res.std_error)
        assert res.std_error < 1.3e-4


def test_two_engines_from_two_host_threads(pkg):
    """One engine = one host thread at a time; two engines on the same GPU from two threads must not
    interfere (own stream, own workspaces, thread-local error string) and give the serial bits."""
    import threading
    opt = pkg.option(**CFG1)
    serial = pkg.Engine(0)
    want = [serial.price_european(opt, 3_000_000 + i, 1234 + i, pkg.CALL) for i in range(6)]
    want_b = serial.price_bullet(pkg.option(N_STEPS=50, B=120.0, P1=5, P2=40), 200_000, 7)
    serial.close()
    got = {}

    def worker(tid):
        eng = pkg.Engine(0)
        out = []
        for rep in range(3):
            for i in range(6):
                out.append(eng.price_european(opt, 3_000_000 + i, 1234 + i, pkg.CALL))
            out.append(eng.price_bullet(pkg.option(N_STEPS=50, B=120.0, P1=5, P2=40), 200_000, 7))
            with pytest.raises(pkg.McbError):
                eng.price_european(pkg.option(S0=-1.0), 10, 1, pkg.CALL)   # errors stay per-thread
        got[tid] = out
        eng.close()

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for tid in range(2):
        out = got[tid]
        for rep in range(3):
            for i in range(6):
                r = out[rep * 7 + i]
                assert r.sum == want[i].sum and r.sumsq == want[i].sumsq
            assert out[rep * 7 + 6].sum == want_b.sum


# ------------------------------------------------------------------ packed keying (optional second keying)
@pytest.mark.parametrize("first,n", [(0, 4096), (3, 1001), (16384 - 2, 7), (5 * 16384 + 1, 40000), ((1 << 34) - 6, 13)])
def test_packed_keying_payoffs_vs_oracle(engine, orc, pkg, first, n):
    """Path p draws normal p & 3 of the cuRAND stream (seed, subsequence p >> 2): per-path payoffs against the
    double-precision restatement, any first path / count (ragged blocks at both ends, a 2^32 boundary of the
    subsequence index).  Same MUFU-based tolerance as the canonical test."""
    for ot, oot in ((pkg.CALL, orc.CALL), (pkg.PUT, orc.PUT)):
        got = engine.european_packed_payoffs(pkg.option(**CFG1), first, n, 1234, ot)
        _, _, want = orc.european_packed(orc.option(**CFG1), first, n, 1234, oot, want_payoffs=True)
        assert np.allclose(got, want, rtol=2e-5, atol=2e-4), float(np.abs(got - want).max())
    # path 4q under packed keying == path q under canonical keying (normal 0 of subsequence q), up to one rounding
    # of the fused multiply (the packed kernel scales the radius before the trig multiply)
    packed = engine.european_packed_payoffs(pkg.option(**CFG1), 0, 4000, 1234, pkg.CALL)[::4]
    canon = engine.european_payoffs(pkg.option(**CFG1), 0, 1000, 1234, pkg.CALL)
    assert np.allclose(packed, canon, rtol=1e-5, atol=1e-4)


def test_packed_keying_tree_and_price(engine, orc, pkg):
    """The packed kernel's reduction: slot t of a chunk takes the chunk's Philox blocks t, t + 256, ... and adds each
    block's four paths in order; then the same block / segment / final trees as everything else (bit-exact against
    the oracle's trees fed with the kernel's own payoffs).  Prices: call and put within 3 SE of the closed form at
    2^26 paths, put-call parity, agreement with the canonical keying within 4 combined SE."""
    n = 3 * pkg.EUROPEAN_CHUNK + 1234
    opt = pkg.option(**CFG1)
    pay = engine.european_packed_payoffs(opt, 0, n, 1234, pkg.CALL)
    res = engine.price_european_packed(opt, n, 1234, pkg.CALL)
    chunk, pps = pkg.EUROPEAN_CHUNK, pkg.EUROPEAN_PATHS_PER_SLOT
    partials = []
    for c in range(4):
        buf = np.zeros(chunk, dtype=np.float32)
        cnt = min(chunk, n - c * chunk)
        buf[:cnt] = pay[c * chunk:c * chunk + cnt]
        # slot t, k-th accumulated path (k = 4 i + j)  <-  chunk-local path 4 (t + 256 i) + j
        slot_major = buf.reshape(pps // 4, 256, 4).transpose(0, 2, 1).reshape(pps, 256)   # [k][t]
        valid = (np.arange(chunk).reshape(pps // 4, 256, 4).transpose(0, 2, 1).reshape(pps, 256) < cnt)
        s = np.zeros(256, np.float32)
        for k in range(pps):
            s = np.where(valid[k], (s + slot_major[k]).astype(np.float32), s)
        # the block tree on 256 slot values = the oracle's chunk tree with one value per slot
        ps, _ = orc.chunk_tree_f32(s, 256, 1)
        partials.append((ps, np.float32(0.0)))
    seg = orc.segment_tree_f64(np.array(partials, dtype=np.float32))
    want_s, _ = orc.final_tree_f64(seg)
    assert res.sum == want_s
    big = 1 << 26
    L = orc.lib()
    call = engine.price_european_packed(opt, big, 1234, pkg.CALL)
    put = engine.price_european_packed(opt, big, 1234, pkg.PUT)
    exact_c = L.orc_bs_call_exact(100.0, 100.0, 1.0, 0.05, 0.2)
    exact_p = L.orc_bs_put_exact(100.0, 100.0, 1.0, 0.05, 0.2)
    assert abs(call.price - exact_c) < 3.0 * call.std_error and abs(put.price - exact_p) < 3.0 * put.std_error
    assert abs((call.price - put.price) - (100.0 - 100.0 * math.exp(-0.05))) < 4.0 * (call.std_error + put.std_error)
    canon = engine.price_european(opt, big, 1234, pkg.CALL)
    assert abs(call.price - canon.price) < 4.0 * math.hypot(call.std_error, canon.std_error)
    assert (call.sum, call.sumsq) == (lambda r: (r.sum, r.sumsq))(engine.price_european_packed(opt, big, 1234, pkg.CALL))
