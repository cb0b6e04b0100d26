#!/usr/bin/env python
"""Small fixed workloads for ncu captures (run plain first, then the same command under ncu).

    python tools/profile_target.py european|packed|trajectory|trajectory_long|bullet|nested|sweep [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
what = sys.argv[1] if len(sys.argv) > 1 else "european"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
CFG = dict(S0=100.0, K=100.0, r=0.05, v=0.2, T=1.0)
eng = pkg.Engine(0)
for _ in range(reps):
    if what == "european":
        r = eng.price_european(pkg.option(**CFG), 1 << 30, 1234, pkg.CALL)
        print(r)
    elif what == "trajectory":
        import torch
        n, steps = 1 << 20, 252
        buf = torch.empty(n * steps, dtype=torch.float32, device="cuda:0")
        eng.simulate_trajectories  # noqa: B018
        eng._lib.mcb_simulate_trajectories(eng._h, pkg.option(N_STEPS=steps, N_PATHS=n, B=120.0, **CFG), 0, n, 1234,
                                           buf.data_ptr(), None, pkg.DEVICE)
        print(float(buf[-1]))
    elif what == "packed":
        print(eng.price_european_packed(pkg.option(**CFG), 1 << 30, 1234, pkg.CALL))
    elif what == "trajectory_long":
        import torch
        n, steps = 1 << 18, 2048
        buf = torch.empty(n * steps, dtype=torch.float32, device="cuda:0")
        cnt = torch.empty(n * steps, dtype=torch.int32, device="cuda:0")
        eng._lib.mcb_simulate_trajectories(eng._h, pkg.option(N_STEPS=steps, N_PATHS=n, B=120.0, **CFG), 0, n, 1234,
                                           buf.data_ptr(), cnt.data_ptr(), pkg.DEVICE)
        print(float(buf[-1]), int(cnt[-1]))
    elif what == "bullet":
        print(eng.price_bullet(pkg.option(N_STEPS=100, N_PATHS=1 << 22, B=120.0, **CFG), 1 << 22, 1234))
    elif what == "nested":
        import torch
        n, steps = 512, 100
        F = torch.empty(n * steps, dtype=torch.float32, device="cuda:0")
        eng.nested_async(pkg.option(N_STEPS=steps, N_PATHS=n, N_PATHS_INNER=4096, B=120.0, **CFG), 0, n, 1234, 1235,
                         pkg.DISCOUNT_COMPAT, F.data_ptr())
        eng.synchronize()
        print(float(F.mean()))
    elif what == "sweep":
        import numpy as np
        K, V = np.meshgrid(np.linspace(60, 140, 32, dtype=np.float32), np.linspace(0.05, 0.8, 32, dtype=np.float32),
                           indexing="ij")
        out = eng.price_sweep(pkg.option(**CFG), K.ravel(), V.ravel(), 1 << 24, 1234, pkg.CALL)
        print(out[0], out[-1])
eng.close()
