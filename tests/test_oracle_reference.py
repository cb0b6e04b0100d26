"""The oracle against the reference's own outputs (golden fixtures + live oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest


def test_option_data_layout(orc, golden_reference, pkg):
    assert golden_reference["sizeof_option_data"] == 48
    assert C.sizeof(orc.OptionData) == 48
    assert C.sizeof(pkg.OptionData) == 48
    assert [f[0] for f in pkg.OptionData._fields_] == ["S0", "T", "K", "r", "v", "B", "P1", "P2", "N_PATHS",
                                                       "N_PATHS_INNER", "N_STEPS", "step"]


def test_black_scholes_restatement_is_bit_exact(orc, golden_reference):
    L = orc.lib()
    for row in golden_reference["black_scholes"]:
        mine = L.orc_bs_call_reference(row["S0"], row["K"], row["T"], row["r"], row["v"])
        assert mine == np.float32(row["call"]), row
    for row in golden_reference["cnd"]:
        assert L.orc_cnd_reference(row["x"]) == np.float32(row["cnd"]), row


def test_closed_forms(orc):
    L = orc.lib()
    assert L.orc_bs_call_reference(100, 100, 1, 0.05, 0.2) == pytest.approx(10.4505768, abs=1e-6)
    assert L.orc_bs_call_reference(100, 100, 1, 0.10, 0.2) == pytest.approx(13.2696915, abs=1e-6)
    assert L.orc_bs_call_exact(100, 100, 1, 0.05, 0.2) == pytest.approx(10.450583572, abs=1e-8)
    assert L.orc_bs_put_exact(100, 100, 1, 0.05, 0.2) == pytest.approx(5.573526022, abs=1e-8)
    # polynomial CND vs erfc: |err| <~ 1e-5 on prices in the sweep range
    for K in np.linspace(60, 140, 9):
        for v in np.linspace(0.05, 0.8, 6):
            a = L.orc_bs_call_reference(100, float(K), 1, 0.05, float(v))
            b = L.orc_bs_call_exact(100, float(K), 1, 0.05, float(v))
            assert abs(a - b) < 5e-5 * max(1.0, b)


def test_config1_european_vs_closed_form(orc):
    """BASELINE configs[0]: 1e6 paths, CPU vs Black-Scholes, within 3 SE; call and put."""
    n = 1_000_000
    o = orc.option(N_PATHS=n)
    L = orc.lib()
    for kind, exact in ((orc.CALL, L.orc_bs_call_exact(100, 100, 1, 0.05, 0.2)),
                        (orc.PUT, L.orc_bs_put_exact(100, 100, 1, 0.05, 0.2))):
        s, q = orc.european(o, 0, n, 1234, kind)
        price = orc.price_from_sum(s, n, o.r, o.T)
        se = orc.std_error(s, q, n, o.r, o.T)
        assert abs(price - exact) < 3.0 * se, (kind, price, exact, se)


def test_oracle_vs_reference_cpu_fixture(orc, golden_reference):
    """The reference's CPU pricers are unseeded: compare statistically (pooled 8 runs)."""
    for run in golden_reference["monte_carlo_unseeded"]:
        ref = np.array(run["prices"])
        if run["kind"] == "vanilla":
            n = 400_000
            o = orc.option(r=run["r"], N_PATHS=n)
            s, q = orc.european(o, 0, n, 1234, orc.CALL)
        else:
            n = 60_000
            o = orc.option(r=run["r"], N_PATHS=n, N_STEPS=run["n_steps"], B=run["B"], P1=run["P1"], P2=run["P2"])
            s, q = orc.bullet(o, 0, n, 1234)
        price = orc.price_from_sum(s, n, o.r, o.T)
        se = orc.std_error(s, q, n, o.r, o.T)
        se_ref = ref.std(ddof=1) / np.sqrt(ref.size)
        assert abs(price - ref.mean()) < 3.5 * np.hypot(se, se_ref), (run["kind"], price, ref.mean(), se, se_ref)


def test_live_reference_cpu_when_built(orc):
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libref_cpu.so not built here")
    R = orc.ref_cpu()
    assert R.ref_sizeof_option_data() == 48
    L = orc.lib()
    rng = np.random.default_rng(3)
    for _ in range(2000):
        a = [np.float32(rng.uniform(50, 150)), np.float32(rng.uniform(50, 150)), np.float32(rng.uniform(0.05, 3)),
             np.float32(rng.uniform(0, 0.15)), np.float32(rng.uniform(0.05, 0.9))]
        assert L.orc_bs_call_reference(*a) == R.ref_black_scholes(*a)
    n = 1 << 20
    o = orc.option(N_PATHS=n)
    ref = np.array([R.ref_vanilla_cpu(C.byref(o)) for _ in range(4)])
    s, q = orc.european(o, 0, n, 1234, orc.CALL)
    price = orc.price_from_sum(s, n, o.r, o.T)
    se = orc.std_error(s, q, n, o.r, o.T)
    assert abs(price - ref.mean()) < 3.5 * se * np.sqrt(1 + 1 / 4)


def test_bullet_restart_and_degenerate_cases(orc):
    n = 2000
    o = orc.option(N_PATHS=n, N_STEPS=40, B=0.0, P1=0, P2=40)  # barrier never hit, window open
    s, q, pay = orc.bullet(o, 0, n, 1234, want_payoffs=True)
    assert s > 0
    # with the window closed from below nothing pays
    o2 = orc.option(N_PATHS=n, N_STEPS=40, B=0.0, P1=1, P2=40)
    s2, _ = orc.bullet(o2, 0, n, 1234)
    assert s2 == 0.0
    # Tk = N_STEPS: zero steps left, payoff of the start price itself
    o3 = orc.option(N_PATHS=n, N_STEPS=40, B=200.0, P1=0, P2=40)
    s3, _, p3 = orc.bullet(o3, 0, 8, 1234, Ik=3, Sk=130.0, Tk=40, want_payoffs=True)
    np.testing.assert_allclose(p3, 30.0, rtol=1e-6)


def test_trajectories_consistent_with_bullet(orc):
    n = 300
    o = orc.option(N_PATHS=n, N_STEPS=37)
    prices, counts = orc.trajectories(o, 5, n, 1234)
    _, _, pay = orc.bullet(o, 5, n, 1234, want_payoffs=True)
    ok = (counts[:, -1] >= o.P1) & (counts[:, -1] <= o.P2)
    expect = np.where(ok, np.maximum(prices[:, -1] - o.K, 0), 0).astype(np.float32)
    np.testing.assert_allclose(pay, expect, rtol=1e-5, atol=1e-5)
    assert (np.diff(counts, axis=1) >= 0).all() and (np.diff(counts, axis=1) <= 1).all()
    assert ((prices < o.B) == (np.diff(np.concatenate([np.zeros((n, 1), np.int32), counts], 1), axis=1) == 1)).mean() > 0.999


def test_nmc_limits(orc):
    L = orc.lib()
    # last step: no inner steps left, F = disc * payoff(S, I) exactly
    o = orc.option(N_PATHS=4, N_PATHS_INNER=8, N_STEPS=12, B=120.0, P1=0, P2=12)
    F, prices, counts = orc.nmc(o, 0, 4, 1234, 1235, orc.DISCOUNT_COMPAT)
    disc = np.exp(-o.r * o.T)
    expect = disc * np.maximum(prices[:, -1].astype(np.float64) - o.K, 0) * ((counts[:, -1] >= 0) & (counts[:, -1] <= 12))
    np.testing.assert_allclose(F[:, -1], expect, rtol=1e-5, atol=1e-6)
    # barrier disabled, window open: F[p,k] -> Black-Scholes C(S[p,k], K, T - t_{k+1}) in CORRECT mode
    o = orc.option(N_PATHS=2, N_PATHS_INNER=4000, N_STEPS=10, B=0.0, P1=0, P2=10)
    F, prices, _ = orc.nmc(o, 0, 2, 1234, 1235, orc.DISCOUNT_CORRECT)
    for p in range(2):
        for k in range(9):
            tau = o.T - (k + 1) * o.step
            bs = L.orc_bs_call_exact(float(prices[p, k]), o.K, tau, o.r, o.v)
            # payoff std <~ 0.35*S over sqrt(4000)
            assert abs(F[p, k] - bs) < 4.0 * 0.35 * prices[p, k] / np.sqrt(4000), (p, k, F[p, k], bs)


def test_sweep_matches_separate_calls(orc):
    n = 5000
    o = orc.option(N_PATHS=n)
    K = np.array([60, 100, 140, 100], dtype=np.float32)
    V = np.array([0.05, 0.2, 0.8, 0.5], dtype=np.float32)
    s, q = orc.sweep(o, K, V, 0, n, 1234)
    for i in range(4):
        oi = orc.option(N_PATHS=n, K=float(K[i]), v=float(V[i]))
        si, qi = orc.european(oi, 0, n, 1234)
        assert s[i] == pytest.approx(si, rel=1e-12) and q[i] == pytest.approx(qi, rel=1e-12)


def test_pregen_path(orc):
    rng = np.random.default_rng(5)
    z = rng.standard_normal((64, 10)).astype(np.float32)
    o = orc.option(N_PATHS=64, N_STEPS=10)
    pay = orc.pregen_european(o, z)
    drift = (o.r - 0.5 * o.v ** 2) * o.step
    st = o.S0 * np.exp((drift + o.v * np.sqrt(o.step) * z.astype(np.float64)).sum(1))
    np.testing.assert_allclose(pay, np.maximum(st - o.K, 0), rtol=1e-5, atol=1e-5)
