"""Times trajectory_kernel (2^20 x 252, device buffer) with CUDA events: python tools/traj_bench.py [steps] [counts]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 252
want_counts = len(sys.argv) > 2 and sys.argv[2] == "1"
n = 1 << 20
eng = pkg.Engine(0)
opt = pkg.option(N_STEPS=steps, N_PATHS=n, B=120.0)
buf = torch.empty(n * steps, dtype=torch.float32, device="cuda")
cnt = torch.empty(n * steps, dtype=torch.int32, device="cuda") if want_counts else None
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    f = lambda: eng.trajectories_async(opt, 0, n, 1234, buf.data_ptr(), cnt.data_ptr() if want_counts else None, st.cuda_stream)
    for _ in range(5): f()
    st.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): f()
    b.record(); st.synchronize()
ms = a.elapsed_time(b) / 50
print(f"layout={os.environ.get('MCB_TRAJ_LAYOUT','default')} steps={steps} counts={want_counts}: {ms*1e3:.1f} us  {4*n*steps/ms/1e6:.1f} GB/s  {n*steps/ms/1e9:.3f} Tsteps/s  chk={float(buf[-1]):.4f}")
