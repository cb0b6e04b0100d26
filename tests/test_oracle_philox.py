"""The oracle's integer stream is pinned against Random123 KATs and cuRAND's own header."""
import ctypes as C

import numpy as np
import pytest


def _words(hexes):
    return np.array([int(x, 16) for x in hexes], dtype=np.uint32)


def test_random123_known_answers(orc, golden_philox):
    for k in golden_philox["kat"]:
        out = orc.philox(_words(k["ctr"]), _words(k["key"]))
        assert [f"{int(x):08x}" for x in out] == k["out"]


def test_stream_vectors_match_curand_fixture(orc, golden_philox):
    assert len(golden_philox["stream"]) >= 250
    for v in golden_philox["stream"]:
        out = orc.stream_block(int(v["seed"]), int(v["subsequence"]), int(v["block"]))
        assert [f"{int(x):08x}" for x in out] == v["out"], v


def test_survey_vectors(orc):
    # SURVEY.md 8(c), seed 1234 / 1235
    h = lambda a: " ".join(f"{int(x):08x}" for x in a)
    assert h(orc.stream_block(1234, 0, 0)) == "2090b348 da7cf0ab 4401906f cbca470e"
    assert h(orc.stream_block(1234, 0, 1)) == "9eeede35 1cbe137c fa277093 147edd50"
    assert h(orc.stream_block(1234, 1, 0)) == "d115a128 52fc7c75 c7f33f17 0f1539db"
    assert h(orc.stream_block(1234, 2, 0)) == "8b438957 ec6a41f7 71067dfe 421d9aa5"
    assert h(orc.stream_block(1234, 1023, 0)) == "01673c65 5e632a22 d9aae41d 3aad5500"
    assert h(orc.stream_block(1234, 1024, 0)) == "7a6635c0 3734acf6 04afeb3c d6f7df44"
    assert h(orc.stream_block(1234, (1 << 20) - 1, 0)) == "9f3dfd7f 21ad8681 c87db7a6 2968af95"
    assert h(orc.stream_block(1234, (1 << 30) - 1, 0)) == "f0e3a6f3 1bef78cf 72be17e8 6d8d879a"
    assert h(orc.stream_block(1234, (1 << 32) + 5, 0)) == "e44a26fe 770cb2fb c4bd06fd 2434977f"
    assert h(orc.stream_block(1235, 0, 0)) == "7131223e 03e95e25 fabbf5c1 c6f16189"


def test_normals_match_curand_host_fixture(orc, golden_philox):
    # cuRAND evaluates Box-Muller in float (logf/sinf/cosf); the oracle in double on the same
    # float uniforms: agreement to float rounding pins word order and the sin/cos pairing.
    for v in golden_philox["curand_normals"]:
        z = orc.stream_normals(int(v["seed"]), int(v["subsequence"]), 16)
        np.testing.assert_allclose(z, np.array(v["normals"]), rtol=0, atol=3e-6)


def test_live_curand_host_when_built(orc):
    if not orc.have_curand_host():
        pytest.skip("oracle/_ref/libcurand_host.so not built here")
    cur = orc.curand_host()
    rng = np.random.default_rng(7)
    u32p = C.POINTER(C.c_uint32)
    for _ in range(2000):
        seed = int(rng.integers(0, 1 << 63))
        sub = int(rng.integers(0, 1 << 63))
        blk = int(rng.integers(0, 1 << 40))
        out = np.zeros(4, dtype=np.uint32)
        cur.curand_host_block(seed, sub, blk, out.ctypes.data_as(u32p))
        assert (out == orc.stream_block(seed, sub, blk)).all()
    # sequential words of one stream: curand() walks blocks 0,1,2... in x,y,z,w order
    w = np.zeros(64, dtype=np.uint32)
    cur.curand_host_words(1234, 42, 0, 64, w.ctypes.data_as(u32p))
    mine = np.concatenate([orc.stream_block(1234, 42, b) for b in range(16)])
    assert (w == mine).all()


def test_uniform_maps_stay_in_range(orc):
    L = orc.lib()
    assert L.orc_uniform_u(0) > 0.0          # log(0) can never happen
    assert L.orc_uniform_u(0xFFFFFFFF) <= 1.0
    assert 0.0 < L.orc_angle_v(0) < 1e-8
    assert L.orc_angle_v(0xFFFFFFFF) <= 6.2831860


def test_normal_moments(orc):
    z = np.concatenate([orc.stream_normals(1234, p, 8) for p in range(20000)])
    n = z.size
    assert abs(z.mean()) < 4.0 / np.sqrt(n)
    assert abs((z ** 2).mean() - 1.0) < 4.0 * np.sqrt(2.0 / n)
    assert abs((z ** 3).mean()) < 4.0 * np.sqrt(15.0 / n)
    assert abs((z ** 4).mean() - 3.0) < 4.0 * np.sqrt(96.0 / n)


def test_packed_keying_is_four_successive_curand_normals(orc, golden_philox):
    """Packed keying: path p uses normal p & 3 of subsequence p >> 2, i.e. what four successive
    curand_normal() calls on the state curand_init(seed, p >> 2, 0) return.  Pinned against the committed cuRAND-header
    fixture (curand_normal sequences of several (seed, subsequence) pairs) through the oracle's packed pricer: the
    payoff of packed path 4 s + j is the payoff of the fixture's normal j of subsequence s."""
    o = orc.option(S0=100.0, K=100.0, r=0.05, v=0.2, T=1.0)
    drift, vol = (0.05 - 0.5 * 0.2 * 0.2) * 1.0, 0.2
    checked = 0
    for v in golden_philox["curand_normals"]:
        seed, sub = int(v["seed"]), int(v["subsequence"])
        if sub >= 1 << 61:
            continue
        _, _, pay = orc.european_packed(o, 4 * sub, 4, seed, orc.CALL, want_payoffs=True)
        want = np.maximum(100.0 * np.exp(drift + vol * np.array(v["normals"][:4])) - 100.0, 0.0)
        np.testing.assert_allclose(pay, want, rtol=0, atol=2e-3)   # fixture normals are float32 (3e-6 apart at most)
        checked += 1
    assert checked >= 2
    # packed path 4 q == canonical path q (normal 0 of subsequence q)
    _, _, a = orc.european_packed(o, 0, 64, 1234, orc.CALL, want_payoffs=True)
    _, _, b = orc.european(o, 0, 16, 1234, orc.CALL, want_payoffs=True)
    assert (a[::4] == b).all()
