"""include/compat: the reference's header names and wrapper_* call surface over libmcb200.so."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "include", "compat")
PKG_DIR = os.path.join(ROOT, "monte-carlo-project-cuda_b200")


def _nvcc(src, exe):
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + COMPAT, src,
                    "-o", exe, "-L" + PKG_DIR, "-lmcb200", "-Xlinker", "-rpath", "-Xlinker", PKG_DIR],
                   check=True, capture_output=True)


def test_example_driver_builds(pkg):
    import __graft_entry__ as entry
    entry._load_build_module().build()
    exe = entry.build_examples()
    assert os.access(exe, os.X_OK)


@pytest.mark.parametrize("driver,symbols", [
    ("hello", ("mcb_price_european", "mcb_price_bullet", "mcb_nested_monte_carlo", "mcb_engine_create")),
    ("testing", ("mcb_simulate_trajectories", "mcb_reduce_blocks", "mcb_generate_normals")),
])
def test_reference_drivers_compile_unchanged(tmp_path, pkg, driver, symbols):
    """the reference's own main()s build against the compat headers + the C-ABI."""
    src = f"/root/reference/{driver}.cu"
    if not os.path.exists(src):
        pytest.skip("/root/reference is only present in the build container")
    import __graft_entry__ as entry
    entry._load_build_module().build()
    exe = str(tmp_path / f"{driver}_ref")
    _nvcc(src, exe)
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
    for name in symbols:
        assert name in syms


def test_compat_closed_form_is_bit_exact(tmp_path, golden_reference):
    """compat/BlackandScholes.hpp against the fixture generated from the unmodified reference."""
    rows = golden_reference["black_scholes"]
    cnd = golden_reference["cnd"]
    src = tmp_path / "bs.cpp"
    lines = ['#include <cstdio>', '#include <cstring>', '#include <cstdint>', '#include "BlackandScholes.hpp"',
             'static unsigned bits(float f){unsigned u; memcpy(&u,&f,4); return u;}', 'int main(){ float c;']
    lit = lambda x: f"{float(np.float32(x))!r}f"   # float literal that round-trips the float32 value
    for r in rows:
        lines.append(f'black_scholes_CPU(c, {lit(r["S0"])}, {lit(r["K"])}, {lit(r["T"])}, {lit(r["r"])}, '
                     f'{lit(r["v"])}); printf("%u\\n", bits(c));')
    for r in cnd:
        lines.append(f'printf("%u\\n", bits(CND({lit(r["x"])})));')
    lines.append('return 0;}')
    src.write_text("\n".join(lines))
    exe = tmp_path / "bs"
    subprocess.run(["g++", "-O2", "-I" + COMPAT, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = np.array([int(x) for x in out], dtype=np.uint32).view(np.float32)
    want = np.array([r["call"] for r in rows] + [r["cnd"] for r in cnd], dtype=np.float32)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()


@pytest.mark.gpu
def test_example_driver_runs_and_matches_the_python_surface(pkg, orc):
    import __graft_entry__ as entry
    exe = entry.build_examples()
    out = subprocess.run([exe, "100000", "64"], capture_output=True, text=True, check=True).stdout
    assert "Average GPU : " in out and "Average GPU bullet option atomic : " in out
    assert "call Black Scholes : " in out and "Average CPU Vanilla Option: " in out
    vals = [float(x) for x in [ln for ln in out.splitlines() if ln.startswith("RESULT")][0].split()[1:]]
    cpu_v, cpu_b, gpu_v, gpu_b, gpu_ba, n1, n2, n3, closed = vals
    assert closed == pytest.approx(13.2696915, abs=2e-6)              # SURVEY section 6, hello.cu parameters
    assert abs(gpu_v - closed) < 0.25 and abs(cpu_v - closed) < 0.25   # 1e5 paths: SE ~ 0.05
    assert gpu_b == gpu_ba and abs(gpu_b - cpu_b) < 0.2
    assert n1 == n2 == n3 and n1 >= 0.0
    # same bits as the Python mirror of the wrappers (same engine, same seeds)
    opt = pkg.option(r=0.1, N_PATHS=100000, N_PATHS_INNER=64, N_STEPS=100)
    assert np.float32(gpu_v) == np.float32(pkg.wrapper_gpu_option_vanilla(opt, 1024, quiet=True))
    assert np.float32(gpu_b) == np.float32(pkg.wrapper_gpu_bullet_option(opt, 1024, quiet=True))


@pytest.mark.gpu
def test_testing_driver_runs(tmp_path, pkg, orc, engine):
    """examples/testing_b200.cu: Simulation facade (pre-generated normals CPU vs GPU, the four
    reductions, outer trajectories) and the CSV format of testing.cu:37-47."""
    import __graft_entry__ as entry
    entry.build_examples()
    exe = os.path.join(ROOT, "build", "testing_b200")
    csv_path = str(tmp_path / "testing.csv")
    out = subprocess.run([exe, csv_path], capture_output=True, text=True, check=True).stdout
    tag = lambda t: [ln.split()[1:] for ln in out.splitlines() if ln.startswith(t)]
    n, worst = tag("PREGEN")[0]
    # same normals; both sides accumulate 100 FP32 log-increments (CPU natural log, GPU log2), so they
    # agree to ~100 * 2^-24 * ln(S) relative: < 2e-2 absolute on payoffs up to a few hundred
    assert int(n) == 1024 and float(worst) < 2e-2
    z = engine.generate_normals(1024 * 100, 1234)
    red = {int(k): np.float32(v) for k, v in tag("REDUCE")}
    first2048 = orc.reduce_sum_f32(z[:2048])
    assert red[3] == red[4] == red[5] == first2048            # block 0 of reduce3/4/5 covers 2*1024 elements
    # reduce6 grid-strides: with one block it covers everything, span by span through the same tree
    assert red[6] == engine.reduce_blocks(z, 1, 2048, strided=True)[0]
    assert abs(float(red[6]) - z.astype(np.float64).sum()) < 0.05
    assert abs(float(tag("HOSTSUM")[0][0]) - z.astype(np.float64).sum()) < 0.5
    # CSV: header, t = 0 row per trajectory, (1+i)*dt timestamps; byte-identical to the ostream rendering
    text = open(csv_path).read()
    assert text == open(csv_path + ".ostream").read()
    lines = text.splitlines()
    assert lines[0] == "time,trajectory,value" and len(lines) == 1 + 20 * 151
    assert lines[1] == "0,0,100" and lines[2].startswith("0.00666667,0,")
    rows = engine.simulate_trajectories(pkg.option(r=0.1, B=0.0, N_STEPS=150, N_PATHS=20, P1=0, P2=0), 0, 20, 555)
    assert float(lines[2].split(",")[2]) == pytest.approx(float(rows[0, 0]), rel=1e-5)
    assert float(lines[-1].split(",")[2]) == pytest.approx(float(rows[-1, -1]), rel=1e-5)


@pytest.mark.gpu
def test_reference_mains_run_on_the_new_engine():
    """The reference's UNMODIFIED hello.cu / testing.cu, compiled in the build container against
    include/compat (build/*_reference_main travel with the snapshot), run on the B200."""
    hello = os.path.join(ROOT, "build", "hello_reference_main")
    testing = os.path.join(ROOT, "build", "testing_reference_main")
    if not (os.path.exists(hello) and os.path.exists(testing)):
        pytest.skip("build/*_reference_main are built only where /root/reference exists")
    out = subprocess.run([hello], capture_output=True, text=True, check=True, timeout=600).stdout
    val = lambda label: float([ln for ln in out.splitlines() if ln.startswith(label)][0].split(":")[1])
    bs = val("call Black Scholes")
    assert bs == pytest.approx(13.2697, abs=1e-4)
    assert abs(val("Average GPU ") - bs) < 0.25 and abs(val("Average CPU Vanilla Option") - bs) < 0.25
    assert val("Average GPU bullet option ") == val("Average GPU bullet option atomic ")
    nmc = [val("Average GPU bullet option nmc " + k) for k in ("one point per block", "one kernel", "optimal")]
    assert nmc[0] == nmc[1] == nmc[2] and nmc[0] > 0
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = subprocess.run([testing], capture_output=True, text=True, check=True, cwd=d, timeout=600).stdout
        assert "Testing reduction: 6" in out and "Total size: 3000" in out
        lines = open(os.path.join(d, "testing.csv")).read().splitlines()
        assert lines[0] == "time,trajectory,value" and len(lines) == 1 + 20 * 151


def test_c_abi_from_plain_c_builds_and_fails_loudly_without_a_gpu(pkg):
    """examples/price_c.c: gcc -std=c99 against include/mcb200.h only.  On a box without a B200 the
    program must report "no engine" and exit 2 -- there is no CPU fallback behind the C-ABI."""
    import torch
    import __graft_entry__ as entry
    entry._load_build_module().build()
    entry.build_examples()
    exe = os.path.join(ROOT, "build", "price_c")
    assert os.access(exe, os.X_OK)
    if not torch.cuda.is_available():
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 2 and "no engine" in out.stderr and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_c_abi_from_plain_c_runs(pkg, engine):
    import __graft_entry__ as entry
    entry.build_examples()
    out = subprocess.run([os.path.join(ROOT, "build", "price_c")], capture_output=True, text=True, check=True).stdout
    vals = [ln for ln in out.splitlines() if ln.startswith("C_ABI")][0].split()[1:]
    rc, bad_rc = int(vals[0]), int(vals[1])
    call, se, put, bullet, first, last = map(float, vals[2:])
    assert rc == 0 and bad_rc == pkg.ERR_INVALID
    opt = pkg.option(N_PATHS=1 << 20, N_STEPS=100, N_PATHS_INNER=1000, step=0.01)
    assert call == engine.price_european(opt, 0, 1234, pkg.CALL).price
    assert put == engine.price_european(opt, 0, 1234, pkg.PUT).price
    assert bullet == engine.price_bullet(opt, 0, 1234).price
    rows = engine.simulate_trajectories(opt, 0, 8, 1234)
    assert np.float32(first) == rows[0, 0] and np.float32(last) == rows[7, 99]
    assert abs((call - put) - (100.0 - 100.0 * np.exp(-0.05))) < 4 * 20.0 / 1024


def test_cmake_build_with_the_reference_targets(tmp_path):
    """CMakeLists.txt: the engine, the examples and -- with -DMCB_REFERENCE_DIR -- the reference's own
    `main` / `test` targets (CMakeLists.txt:20-21 of the reference) configure and build for sm_100a."""
    import shutil
    if not (shutil.which("cmake") and shutil.which("ninja")):
        pytest.skip("cmake / ninja not on PATH")
    args = ["cmake", "-S", ROOT, "-B", str(tmp_path), "-G", "Ninja"]
    have_ref = os.path.exists("/root/reference/hello.cu")
    if have_ref:
        args.append("-DMCB_REFERENCE_DIR=/root/reference")
    subprocess.run(args, check=True, capture_output=True)
    subprocess.run(["cmake", "--build", str(tmp_path), "-j", "4"], check=True, capture_output=True)
    for name in ["libmcb200.so", "hello_b200", "testing_b200"] + (["main", "test"] if have_ref else []):
        assert os.path.exists(tmp_path / name), name
