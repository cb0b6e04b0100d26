#!/usr/bin/env python
"""Every kernel at small, ragged sizes -- the target for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as entry

pkg = entry.load_package()
eng = pkg.Engine(0)
o = pkg.option()
for n in (1, 257, 16385, 100001):
    eng.price_european(o, n, 1234, pkg.CALL)
    eng.price_european(o, n, 1234, pkg.PUT)
eng.european_payoffs(o, 16000, 1000)
eng.european_chunk_partials(o, 40000)
ob = pkg.option(N_STEPS=37, N_PATHS=3001, B=120.0, P1=1, P2=30)
eng.price_bullet(ob, 3001, 1234)
eng.bullet_payoffs(ob, 5, 1031, 1234, Ik=2, Sk=101.0, Tk=4)
for steps, paths in ((252, 61), (100, 33), (7, 5), (1, 3), (256, 24), (300, 9), (128, 40)):
    t = pkg.option(N_STEPS=steps, N_PATHS=paths, B=120.0)
    eng.simulate_trajectories(t, 3, paths, 1234)
    eng.simulate_trajectories(t, 3, paths, 1234, want_counts=True)
nm = pkg.option(N_STEPS=13, N_PATHS=5, N_PATHS_INNER=300, B=120.0, P1=1, P2=10)
eng.nested_monte_carlo(nm, 2, 5, 1234, 1235, pkg.DISCOUNT_CORRECT)
k = np.linspace(80, 120, 7, dtype=np.float32)
v = np.linspace(0.1, 0.5, 7, dtype=np.float32)
eng.price_sweep(o, k, v, 20001, 1234, pkg.CALL)
eng.price_sweep(o, k, v, 20001, 1234, pkg.PUT)
eng.reduce_sum(np.ones(1000, np.float32))
eng.reduce_blocks(np.ones(5000, np.float32), 3, 512, strided=True)
eng.reduce_blocks(np.ones(5000, np.float32), 3, 512, strided=False)
eng.generate_normals(1001)
eng.price_from_normals(pkg.option(N_STEPS=5), np.zeros((77, 5), np.float32))
eng.philox_blocks(1, [0, 1, 2], [0, 1, 2])
eng.philox_blocks(1, [0, 1, 2], [0, 1, 2], library=True)
eng.stream_normals(1, 2, 1003, n0=1)
eng.close()
print("sanitize target ok")
