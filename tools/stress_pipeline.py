"""Stress of the job pipeline over CUDA-IPC mailboxes (run under torchrun, one process per GPU):
thousands of small and mid-size jobs submitted back to back (ring wrap-arounds, acks, both pricing streams, empty
shards), every result compared bit for bit with a single-GPU engine of the same process.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/stress_pipeline.py [jobs]
Run with plain `python` it stresses ONE engine handle over every GPU of the process instead (launcher threads, events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib

import torch
import torch.distributed as dist

import __graft_entry__ as entry

pkg = entry.load_package()
jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
INPROC = "RANK" not in os.environ                  # plain `python`: ONE engine over every GPU of the process
if INPROC:
    devices = [int(d) for d in os.environ["MCB_STRESS_DEVICES"].split(",")] if os.environ.get("MCB_STRESS_DEVICES") \
        else list(range(torch.cuda.device_count()))        # e.g. MCB_STRESS_DEVICES=0,0,0: three shards on one GPU
    rank, world, local = 0, len(devices), 0
    eng = pkg.Engine(devices)
    solo = pkg.Engine(0)
else:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sharded = importlib.import_module(entry.PKG_NAME + ".sharded")
    eng = pkg.Engine(local)
    solo = pkg.Engine(local)
    pricer = sharded.ShardedPricer(eng, transport="peer")
sizes = [1, 100_000, 16384 * 63 + 5, 16384 * 200 + 999, 1 << 22, 16384 * 4100 + 3]
bad = 0
pending = []
for i in range(jobs):
    n = sizes[i % len(sizes)]
    typ = pkg.PUT if i % 3 == 0 else pkg.CALL
    pending.append((eng.european_submit(pkg.option(K=90.0 + i % 20), n, 1234 + i % 7, typ), n, 90.0 + i % 20, 1234 + i % 7, typ))
    if len(pending) == pkg.PIPELINE_DEPTH:          # keep the ring full: collect the oldest only
        t, n0, k0, s0, ty0 = pending.pop(0)
        got = eng.european_collect(t)
        if i % 50 == 0:                              # the single-GPU comparison is the slow part: sample it
            want = solo.price_european(pkg.option(K=k0), n0, s0, ty0)
            bad += (got.sum, got.sumsq, got.price) != (want.sum, want.sumsq, want.price)
for t, n0, k0, s0, ty0 in pending:
    got = eng.european_collect(t)
    want = solo.price_european(pkg.option(K=k0), n0, s0, ty0)
    bad += (got.sum, got.sumsq, got.price) != (want.sum, want.sumsq, want.price)
flag = torch.tensor([bad, eng.peer_timeouts()], dtype=torch.int64, device="cuda")
if not INPROC:
    dist.all_reduce(flag)
if rank == 0:
    shape = "one engine over %d shards" % world if INPROC else "%d ranks" % world
    print(f"stress: {jobs} jobs on {shape}, mismatches {int(flag[0])}, timeouts {int(flag[1])}")
eng.close()
solo.close()
if not INPROC:
    dist.destroy_process_group()
sys.exit(1 if int(flag[0]) or int(flag[1]) else 0)
