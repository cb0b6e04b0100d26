"""CPU-side checks: reduction-tree restatement, shard arithmetic, C-ABI surface."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_tree_against_exact_sum(orc):
    rng = np.random.default_rng(1)
    for n_valid, pps in ((16384, 64), (16000, 64), (1, 64), (255, 64), (1024, 4), (1000, 4), (0, 4)):
        pay = rng.gamma(2.0, 5.0, size=256 * pps).astype(np.float32)
        s, q = orc.chunk_tree_f32(pay, n_valid, pps)
        exact = pay[:n_valid].astype(np.float64).sum()
        exact_q = (pay[:n_valid].astype(np.float64) ** 2).sum()
        assert abs(float(s) - exact) <= 2e-6 * max(exact, 1e-30)
        assert abs(float(q) - exact_q) <= 2e-6 * max(exact_q, 1e-30)


def test_chunk_tree_is_order_defined(orc):
    # the tree is a function of the slot index: permuting payoffs WITHIN a slot's sequence
    # changes bits, the documented order does not.
    pay = (np.arange(16384, dtype=np.float32) % 97) * np.float32(0.37)
    a = orc.chunk_tree_f32(pay, 16384, 64)
    b = orc.chunk_tree_f32(pay.copy(), 16384, 64)
    assert a == b
    # hand-rolled restatement in numpy float32
    x = pay.reshape(64, 256)
    s = np.zeros(256, np.float32)
    for i in range(64):
        s = (s + x[i]).astype(np.float32)
    w = s.reshape(8, 32)
    for off in (16, 8, 4, 2, 1):
        w[:, :off] = w[:, :off] + w[:, off:2 * off]
    y = w[:, 0].copy()
    for off in (4, 2, 1):
        y[:off] = y[:off] + y[off:2 * off]
    assert np.float32(y[0]) == a[0]


def test_segment_ranges_partition(orc, pkg):
    for n_chunks in (0, 1, 7, 63, 64, 65, 1000, 65536, 65537):
        prev = 0
        for s in range(orc.SEGMENTS):
            lo, hi = orc.segment_range(n_chunks, s)
            assert lo == prev and hi >= lo
            prev = hi
        assert prev == n_chunks
        for world in (1, 2, 3, 4, 5, 6, 7, 8, 16, 64):
            cover = 0
            seg_prev = 0
            for rank in range(world):
                seg_lo, seg_hi, c_lo, c_hi = pkg.segment_span(rank, world, n_chunks)
                assert seg_lo == seg_prev and c_lo == cover
                assert (c_lo, c_hi) == (orc.segment_range(n_chunks, seg_lo)[0] if seg_lo < 64 else n_chunks,
                                        orc.segment_range(n_chunks, seg_hi - 1)[1] if seg_hi > seg_lo else c_lo)
                seg_prev, cover = seg_hi, c_hi
            assert seg_prev == 64 and cover == n_chunks
    with pytest.raises(ValueError):
        pkg.segment_span(2, 2, 10)


def test_path_span_partition(pkg):
    for n in (0, 1, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [pkg.path_span(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_segment_and_final_tree(orc):
    rng = np.random.default_rng(2)
    for n_chunks in (1, 5, 64, 100, 4097):
        cp = rng.gamma(2.0, 1e5, size=(n_chunks, 2)).astype(np.float32)
        seg = orc.segment_tree_f64(cp)
        s, q = orc.final_tree_f64(seg)
        assert s == pytest.approx(cp[:, 0].astype(np.float64).sum(), rel=1e-14)
        assert q == pytest.approx(cp[:, 1].astype(np.float64).sum(), rel=1e-14)
        # rank-sharded evaluation: owned segments + zeros elsewhere, summed in any order, same bits
        for world in (2, 3, 8):
            total = np.zeros_like(seg)
            for rank in reversed(range(world)):
                lo, hi = rank * 64 // world, (rank + 1) * 64 // world
                part = np.zeros_like(seg)
                part[lo:hi] = seg[lo:hi]
                total = total + part
            assert (total == seg).all()


def test_reduce_sum_restatement(orc):
    rng = np.random.default_rng(4)
    x = rng.standard_normal(102400).astype(np.float32)
    assert abs(float(orc.reduce_sum_f32(x)) - x.astype(np.float64).sum()) < 1e-2
    assert orc.reduce_sum_f32(np.zeros(0, np.float32)) == 0.0


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-z0-9_]+)\s*\(", text)))


def test_c_abi_exports_every_declared_symbol(pkg):
    """The library loads without a GPU and exports exactly what include/mcb200.h declares."""
    import __graft_entry__ as entry
    entry._load_build_module().build()
    lib = pkg.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(pkg.SIGNATURES), set(declared) ^ set(pkg.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r"\bT (mcb_[a-z0-9_]+)", out.stdout))
    assert set(declared) <= exported
    assert lib.mcb_version() == 2


def test_no_device_is_a_loud_error(pkg):
    """Without a GPU the product must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.McbError) as ei:
        pkg.Engine(0)
    assert ei.value.status == pkg.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)
    assert pkg.wrapper_gpu_option_vanilla(pkg.option(), 1024, quiet=True) == -1.0


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "monte-carlo-project-cuda_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert not any("#include" in ln and "oracle" in ln for ln in text.splitlines()), f
                assert "libmc_oracle" not in text, f
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f)).read()
            assert not any("#include" in ln and "oracle" in ln for ln in text.splitlines()), f


def test_python_constants_match_the_header(pkg):
    """The binding's constants are the header's macros / enumerators (geometry of the reduction, pipeline depths,
    status codes, kernel ids): a drift here would silently mis-size buffers or mis-read statuses."""
    text = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    macro = lambda name: int(re.search(r"#define\s+%s\s+(\d+)" % name, text).group(1))
    assert macro("MCB_SLOTS") == pkg.SLOTS and macro("MCB_SEGMENTS") == pkg.SEGMENTS
    assert macro("MCB_EUROPEAN_PATHS_PER_SLOT") == pkg.EUROPEAN_PATHS_PER_SLOT
    assert macro("MCB_BULLET_PATHS_PER_SLOT") == pkg.BULLET_PATHS_PER_SLOT
    assert macro("MCB_PIPELINE_DEPTH") == pkg.PIPELINE_DEPTH and macro("MCB_RESULT_RING") == pkg.RESULT_RING
    assert macro("MCB_MAX_PEERS") == pkg.MAX_PEERS and macro("MCB_IPC_HANDLE_BYTES") == 64
    enum = lambda name: int(re.search(r"\b%s\s*=\s*(\d+)" % name, text).group(1))
    for name, value in (("MCB_OK", pkg.OK), ("MCB_ERR_INVALID", pkg.ERR_INVALID), ("MCB_ERR_CUDA", pkg.ERR_CUDA),
                        ("MCB_ERR_NO_DEVICE", pkg.ERR_NO_DEVICE), ("MCB_ERR_NOMEM", pkg.ERR_NOMEM),
                        ("MCB_ERR_TIMEOUT", pkg.ERR_TIMEOUT), ("MCB_KERNEL_EUROPEAN", pkg.KERNEL_EUROPEAN),
                        ("MCB_KERNEL_SWEEP", pkg.KERNEL_SWEEP), ("MCB_KERNEL_EUROPEAN_PACKED", pkg.KERNEL_EUROPEAN_PACKED),
                        ("MCB_CALL", pkg.CALL), ("MCB_PUT", pkg.PUT), ("MCB_HOST", pkg.HOST), ("MCB_DEVICE", pkg.DEVICE)):
        assert enum(name) == value, name
    assert pkg.Result.__dict__ is not None and __import__("ctypes").sizeof(pkg.Result) == 40
    assert __import__("ctypes").sizeof(pkg.OptionData) == 48


def test_small_job_segment_closed_form():
    """european_small_job_kernel (pricing_kernels.cuh, small_job_tail_one_gpu) finds the segment of chunk c of an
    n-chunk job, n <= 64, as (64 (c + 1) + n - 1) / n - 1 in 32-bit arithmetic instead of searching the segment
    bounds floor(n s / 64) <= c < floor(n (s + 1) / 64) the other kernels and the oracle use: same answer for
    every (n, c), every such segment holds exactly that one chunk, and a one-chunk job lives in segment 63."""
    for n in range(1, 65):
        owners = {}
        for s in range(64):
            for c in range((n * s) // 64, (n * (s + 1)) // 64):
                assert c not in owners
                owners[c] = s
        assert sorted(owners) == list(range(n))
        for c in range(n):
            assert (64 * (c + 1) + n - 1) // n - 1 == owners[c], (n, c)
            assert (n * (owners[c] + 1)) // 64 - (n * owners[c]) // 64 == 1
    assert (64 * 1 + 1 - 1) // 1 - 1 == 63


def test_result_slot_words_round_trip():
    """The flag-in-data result slot (HostSlot, pricing_kernels.cuh / host_slot_read, mcb200.cu): five 8-byte fields travel
    as ten (half | tag << 32) words; a slot is complete only when all ten carry the ticket's tag, and the value an
    expired slot holds can never match it."""
    import struct
    fields = [struct.unpack("<Q", struct.pack("<d", x))[0] for x in (10.45, 0.0143, 1.1e10, float("nan"))] + [1 << 40]
    for ticket in (1, 7, 0xffffffff, 1 << 32, (1 << 32) + 5):
        tag = ticket & 0xffffffff
        words = []
        for f in fields:
            words += [(tag << 32) | (f & 0xffffffff), (tag << 32) | (f >> 32)]
        never = ((~ticket) & 0xffffffff) << 32
        assert all((never >> 32) != (w >> 32) for w in words)
        for missing in range(10):                          # any word still expired: not complete
            partial = list(words)
            partial[missing] = never
            assert not all((w >> 32) == tag for w in partial)
        back = [(words[2 * k] & 0xffffffff) | ((words[2 * k + 1] & 0xffffffff) << 32) for k in range(5)]
        assert back == fields
