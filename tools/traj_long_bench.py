"""Long trajectory rows (more than one 256-step pass): general kernel (MCB_TRAJ_LONG=0), the shipped choice (1: whole-row
staging while it fits, pass-wise staging beyond) and pass-wise staging forced (2); checks that all three write the same bits.
    python tools/traj_long_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry

pkg = entry.load_package()
eng = pkg.Engine(0)
st = torch.cuda.Stream()


def run(opt, n, buf, cnt, reps=10):
    with torch.cuda.stream(st):
        f = lambda: eng.trajectories_async(opt, 0, n, 1234, buf.data_ptr(), cnt.data_ptr() if cnt is not None else None,
                                           st.cuda_stream)
        for _ in range(2):
            f()
        st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            f()
        b.record()
        st.synchronize()
    return a.elapsed_time(b) / reps


for steps in (512, 1024, 1536, 2048, 4096):
    n = (1 << 20) if steps <= 1024 else (1 << 18)
    opt = pkg.option(N_STEPS=steps, N_PATHS=n, B=120.0)
    for want_counts in (False, True):
        ref = None
        for mode in (0, 1, 2):
            os.environ["MCB_TRAJ_LONG"] = str(mode)
            buf = torch.full((n * steps,), float("nan"), dtype=torch.float32, device="cuda")
            cnt = torch.full((n * steps,), -1, dtype=torch.int32, device="cuda") if want_counts else None
            ms = run(opt, n, buf, cnt)
            per = 8 if want_counts else 4
            same = ""
            if ref is None:
                ref = (buf.clone(), cnt.clone() if cnt is not None else None)
            else:
                ok = bool((buf.view(torch.int32) == ref[0].view(torch.int32)).all())
                if cnt is not None:
                    ok = ok and bool((cnt == ref[1]).all())
                same = "bits==general" if ok else "BITS DIFFER"
            print(f"steps={steps} rows=2^{n.bit_length()-1} counts={int(want_counts)} mode={mode}: {ms*1e3:8.1f} us  "
                  f"{per*n*steps/ms/1e6:7.1f} GB/s  {same}", flush=True)
            del buf, cnt
        del ref
        torch.cuda.empty_cache()
os.environ.pop("MCB_TRAJ_LONG", None)
