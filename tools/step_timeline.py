#!/usr/bin/env python
"""Per-op timeline of one decode step with CUDA events on the launching stream (works under torchrun: every rank runs, rank 0
prints). Eager launches, PDL off, one event pair per C-ABI call: the table is the per-kernel launch list of the step at this
parallelism -- ncu cannot attach to a multi-rank run. Usage: [torchrun ...] tools/step_timeline.py [--model ... --batch ... --ctx ...]"""
import argparse
import collections
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtp_llm_b200 import ops  # noqa: E402
from rtp_llm_b200.decode_step import LLAMA3_8B, QWEN2_72B, DecodeStep  # noqa: E402
from rtp_llm_b200.tp import make_comm  # noqa: E402
import dataclasses  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--quant", default="int4")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=2048)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--comm", default="peer")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = make_comm(dev, kind=a.comm)
    base = {"llama3-8b": LLAMA3_8B, "qwen2-72b": QWEN2_72B}[a.model]
    cfg = dataclasses.replace(base, quant=a.quant, layers=a.layers or base.layers)
    m = DecodeStep(cfg, a.batch, a.ctx, dev, tp_rank=rank, tp_size=world, comm=comm, pdl=False)
    events = []

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def inner(*args, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*args, **kw)
            e.record()
            shape = ""
            if name == "wo_gemm":
                shape = f" K{args[1].K} N{args[1].N}"
            events.append((label + shape, s, e))
            return r
        setattr(obj, name, inner)
    for n in ("convert_block_table", "embedding", "add_rmsnorm", "wo_gemm", "rope_append", "paged_decode_attn", "argmax", "silu_and_mul"):
        wrap(ops, n, n)
    if comm is not None:
        for n in ("all_reduce", "all_reduce_norm", "argmax", "all_gather"):
            if hasattr(comm, n):
                wrap(comm, n, "comm." + n)
    for _ in range(3):
        m.step()
    torch.cuda.synchronize()
    events.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    REP = 5
    t0.record()
    for _ in range(REP):
        m.step()
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        agg = collections.OrderedDict()
        for label, s, e in events:
            d = agg.setdefault(label, [0, 0.0])
            d[0] += 1
            d[1] += s.elapsed_time(e) * 1e3
        total = sum(v[1] for v in agg.values()) / REP
        wall = t0.elapsed_time(t1) * 1e3 / REP
        print(f"# {cfg.name} {a.quant} B{a.batch} ctx{a.ctx} tp{world} comm={a.comm}: eager step, PDL off, CUDA events per call, mean of {REP} steps")
        print(f"# sum of op times {total:.1f} us; eager step wall {wall:.1f} us (host launch gaps included); launches/step {sum(v[0] for v in agg.values()) // REP}")
        print("op,calls_per_step,avg_us,total_us,share")
        for label, (n, us) in agg.items():
            print(f"{label},{n // REP},{us / n:.2f},{us / REP:.1f},{us / REP / total:.3f}")
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
