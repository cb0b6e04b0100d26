"""Times the trajectory kernels (2^20 rows, device buffer) with CUDA events and checks that every launcher
variant writes the same bits:  python tools/traj_bench.py [steps ...]   (default 252)

The slab kernel's FAST flag (hoisted Philox products + packed FP32x2) is forced on / off through the
launcher's MCB_TRAJ_FAST knob (csrc/mcb200.cu launch_trajectory); unset, the launcher uses it when counts or
log2 prices are stored next to the prices."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry

pkg = entry.load_package()
steps_list = [int(a) for a in sys.argv[1:]] or [252]
n = 1 << 20
eng = pkg.Engine(0)
st = torch.cuda.Stream()


def run(opt, buf, cnt, reps=50):
    with torch.cuda.stream(st):
        f = lambda: eng.trajectories_async(opt, 0, n, 1234, buf.data_ptr(), cnt.data_ptr() if cnt is not None else None,
                                           st.cuda_stream)
        for _ in range(5):
            f()
        st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            f()
        b.record()
        st.synchronize()
    return a.elapsed_time(b) / reps


for steps in steps_list:
    opt = pkg.option(N_STEPS=steps, N_PATHS=n, B=120.0)
    for want_counts in (False, True):
        ref = None
        for fast in (0, 1):   # the slab kernel's FAST flag: hoisted Philox products + packed FP32x2
            os.environ["MCB_TRAJ_FAST"] = str(fast)
            buf = torch.full((n * steps,), float("nan"), dtype=torch.float32, device="cuda")
            cnt = torch.full((n * steps,), -1, dtype=torch.int32, device="cuda") if want_counts else None
            ms = run(opt, buf, cnt)
            per = 8 if want_counts else 4
            same = ""
            if ref is None:
                ref = (buf.clone(), cnt.clone() if cnt is not None else None)
            else:
                ok = bool((buf.view(torch.int32) == ref[0].view(torch.int32)).all())
                if cnt is not None:
                    ok = ok and bool((cnt == ref[1]).all())
                same = "bits==plain" if ok else "BITS DIFFER"
            print(f"steps={steps} counts={int(want_counts)} fast={fast}: {ms*1e3:7.1f} us  "
                  f"{per*n*steps/ms/1e6:7.1f} GB/s  {same}", flush=True)
            del buf, cnt
os.environ.pop("MCB_TRAJ_FAST", None)
