"""Regenerates the golden fixtures in this directory.  Run in the BUILD container only
(needs /root/reference and nvcc):   python tests/golden/make_golden.py

* philox_vectors.json  -- Random123 known-answer tests (published with the Philox paper,
  quoted in SURVEY.md 8c) + stream vectors and normals produced by cuRAND's OWN
  Philox4x32-10 header compiled for the host (oracle/curand_host.cu).
* reference_cpu.json   -- outputs of the UNMODIFIED reference (oracle/ref_harness.cu compiled
  against /root/reference/inc): black_scholes_CPU, CND, sizeof(OptionData), and a few runs of
  its (unseeded, std::random_device) CPU Monte Carlo pricers for statistical comparison.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def hexwords(a):
    return [f"{int(x):08x}" for x in a]


def main():
    oracle.build(ref=True)
    cur = oracle.curand_host()
    ref = oracle.ref_cpu()
    rng = np.random.default_rng(20261018)
    u32p = C.POINTER(C.c_uint32)

    kats = [
        {"ctr": ["00000000"] * 4, "key": ["00000000"] * 2,
         "out": ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]},
        {"ctr": ["ffffffff"] * 4, "key": ["ffffffff"] * 2,
         "out": ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]},
        {"ctr": ["243f6a88", "85a308d3", "13198a2e", "03707344"], "key": ["a4093822", "299f31d0"],
         "out": ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]},
    ]
    # the same three through cuRAND's curand_Philox4x32_10 on the host must agree
    for k in kats:
        c = np.array([int(x, 16) for x in k["ctr"]], dtype=np.uint32)
        key = np.array([int(x, 16) for x in k["key"]], dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        cur.curand_host_philox(c.ctypes.data_as(u32p), key.ctypes.data_as(u32p), out.ctypes.data_as(u32p))
        assert hexwords(out) == k["out"], (hexwords(out), k)

    cases = [(1234, 0, 0), (1234, 0, 1), (1234, 1, 0), (1234, 1, 1), (1234, 2, 0), (1234, 1023, 0),
             (1234, 1024, 0), (1234, (1 << 20) - 1, 0), (1234, (1 << 30) - 1, 0), (1234, (1 << 32) + 5, 0),
             (1235, 0, 0), (1234, (1 << 32) - 1, 0), (1234, 1 << 32, 7), (0, 0, 0),
             ((1 << 64) - 1, (1 << 64) - 1, (1 << 40) + 3)]
    for _ in range(240):
        seed = int(rng.integers(0, 1 << 63)) if rng.random() < 0.5 else int(rng.integers(0, 5000))
        sub = int(rng.integers(0, 1 << 62)) if rng.random() < 0.4 else int(rng.integers(0, 1 << 31))
        blk = int(rng.integers(0, 1 << 34)) if rng.random() < 0.2 else int(rng.integers(0, 64))
        cases.append((seed, sub, blk))
    stream = []
    for seed, sub, blk in cases:
        out = np.zeros(4, dtype=np.uint32)
        cur.curand_host_block(seed, sub, blk, out.ctypes.data_as(u32p))
        stream.append({"seed": str(seed), "subsequence": str(sub), "block": str(blk), "out": hexwords(out)})
    # SURVEY.md 8(c) quotes these; keep the generator honest
    assert stream[0]["out"] == ["2090b348", "da7cf0ab", "4401906f", "cbca470e"]
    assert stream[9]["out"] == ["e44a26fe", "770cb2fb", "c4bd06fd", "2434977f"]
    assert stream[10]["out"] == ["7131223e", "03e95e25", "fabbf5c1", "c6f16189"]

    normals = []
    for seed, sub in [(1234, 0), (1234, 7), (1235, 123456789), (99, (1 << 33) + 1)]:
        z = np.zeros(16, dtype=np.float32)
        cur.curand_host_normals(seed, sub, 16, z.ctypes.data_as(C.POINTER(C.c_float)))
        normals.append({"seed": str(seed), "subsequence": str(sub), "normals": [float(x) for x in z]})

    with open(os.path.join(HERE, "philox_vectors.json"), "w") as f:
        json.dump({"source": "Random123 KATs + cuRAND 10.3.10 (CUDA 12.9) curand_philox4x32_x.h compiled for host",
                   "kat": kats, "stream": stream, "curand_normals": normals}, f, indent=1)

    bs = []
    grid = [(100, 100, 1, 0.05, 0.2), (100, 100, 1, 0.1, 0.2), (100, 60, 1, 0.05, 0.05), (100, 140, 1, 0.05, 0.8),
            (100, 120, 0.5, 0.02, 0.35), (80, 100, 2.0, 0.0, 0.15), (120, 100, 0.25, 0.07, 0.6)]
    for _ in range(60):
        grid.append((float(np.float32(rng.uniform(40, 160))), float(np.float32(rng.uniform(40, 160))),
                     float(np.float32(rng.uniform(0.05, 3.0))), float(np.float32(rng.uniform(0.0, 0.15))),
                     float(np.float32(rng.uniform(0.05, 0.9)))))
    for S0, K, T, r, v in grid:
        bs.append({"S0": S0, "K": K, "T": T, "r": r, "v": v,
                   "call": float(ref.ref_black_scholes(S0, K, T, r, v))})
    cnd = [{"x": float(np.float32(x)), "cnd": float(ref.ref_cnd(float(np.float32(x))))}
           for x in np.linspace(-6, 6, 49)]

    mc = []
    o = oracle.option(N_PATHS=1_000_000)
    mc.append({"kind": "vanilla", "r": 0.05, "n_paths": 1_000_000,
               "prices": [float(ref.ref_vanilla_cpu(C.byref(o))) for _ in range(8)]})
    ob = oracle.option(N_PATHS=100_000, N_STEPS=100)
    mc.append({"kind": "bullet", "r": 0.05, "n_paths": 100_000, "n_steps": 100, "B": 120.0, "P1": 10, "P2": 50,
               "prices": [float(ref.ref_bullet_cpu(C.byref(ob))) for _ in range(8)]})
    oh = oracle.option(r=0.1, N_PATHS=100_000, N_STEPS=100)  # hello.cu:5-17
    mc.append({"kind": "bullet", "r": 0.1, "n_paths": 100_000, "n_steps": 100, "B": 120.0, "P1": 10, "P2": 50,
               "prices": [float(ref.ref_bullet_cpu(C.byref(oh))) for _ in range(8)]})

    with open(os.path.join(HERE, "reference_cpu.json"), "w") as f:
        json.dump({"source": "unmodified reference compiled from /root/reference/inc via oracle/ref_harness.cu",
                   "sizeof_option_data": int(ref.ref_sizeof_option_data()),
                   "black_scholes": bs, "cnd": cnd, "monte_carlo_unseeded": mc}, f, indent=1)
    print("wrote philox_vectors.json, reference_cpu.json")


if __name__ == "__main__":
    main()
