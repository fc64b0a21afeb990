#!/usr/bin/env python
"""Peer-memory TP collectives against torch references, on real GPUs (launch with torchrun --nproc-per-node N; used by
tests/test_gpu_tp.py when >= 2 GPUs are visible): one-shot and two-shot all-reduce, the fused all-reduce + residual +
RMSNorm, vocab-parallel argmax, and a CUDA graph holding an ODD number of all-reduces replayed several times (slot parity
comes from the device-side call counter)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtp_llm_b200 import ops  # noqa: E402
from rtp_llm_b200.tp import make_comm  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    comm = make_comm(dev, kind="peer")
    ok = True

    def report(name, good, info=""):
        nonlocal ok
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        good = bool(flag.item())
        ok &= good
        if rank == 0:
            print(f"[{'PASS' if good else 'FAIL'}] tp{world} {name} {info}", flush=True)

    def ref_sum(t):
        """Sum in rank order in fp32, rounded once -- what the kernels compute (identical bits on every rank)."""
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        acc = torch.zeros_like(t, dtype=torch.float32)
        for p in parts:
            acc += p.float()
        return acc.to(t.dtype)

    for dtype in (torch.float16, torch.bfloat16):
        for rows, hidden in ((32, 4096), (5, 1024), (1, 8192)):
            g = torch.Generator(device=dev).manual_seed(100 * rank + rows)
            t = torch.randn(rows, hidden, generator=g, device=dev).to(dtype)
            exp = ref_sum(t)
            for two_shot in (0, 1):
                os.environ["B200_AR_TWOSHOT"] = str(two_shot)
                x = t.clone()
                comm.all_reduce(x)
                torch.cuda.synchronize()
                report(f"all_reduce {'two' if two_shot else 'one'}-shot {dtype} [{rows},{hidden}]", torch.equal(x, exp))
            os.environ.pop("B200_AR_TWOSHOT", None)
            # fused all-reduce + residual + rmsnorm vs the unfused sequence through our own kernels
            gamma = (1 + 0.1 * torch.randn(hidden, generator=torch.Generator(device=dev).manual_seed(7), device=dev)).to(dtype)
            resid = torch.randn(rows, hidden, generator=torch.Generator(device=dev).manual_seed(9), device=dev).to(dtype)
            r1, r2 = resid.clone(), resid.clone()
            y1 = torch.empty_like(t)
            fused = comm.all_reduce_norm(t.clone(), r1, gamma, 1e-5, y1)
            x2 = t.clone()
            comm.all_reduce(x2)
            y2 = ops.add_rmsnorm(x2, r2, gamma, 1e-5)
            torch.cuda.synchronize()
            if fused:
                err = (y1.float() - y2.float()).abs().max().item()
                report(f"all_reduce_norm {dtype} [{rows},{hidden}]", torch.equal(r1, r2) and err <= 2e-2, f"max |dy| {err:.3g}")
    # GEMM + reduce-scatter in one kernel (b200_wo_gemm_rs) followed by gather + residual + norm (b200_peer_gather_norm) against
    # the unfused sequence b200_wo_gemm -> b200_peer_allreduce_norm: residual and y must be BIT-identical (same rounded partials,
    # same summation order); covers the direct epilogue (one k-split), the cluster split-K merge, all weight formats, a graph replay
    import numpy as np
    for dtype in (torch.float16, torch.bfloat16):
        for (fmt, B, K, N) in (("int4", 32, 512, 4096), ("int4", 32, 2048, 4096), ("int4", 5, 1024, 1024), ("int8", 17, 512, 2048),
                               ("int8g", 32, 256, 4096), ("f16", 8, 256, 1024), ("int4", 64, 1792, 4096), ("int4", 100, 512, 2048)):
            if not comm.gemm_rs_supported(B, N):
                continue
            gw = torch.Generator(device=dev).manual_seed(1000 * rank + K + N)
            x = (torch.randn(B, K, generator=gw, device=dev) * 0.5).to(dtype)
            if fmt == "int4":
                qp = torch.randint(0, 256, (K, N // 2), generator=gw, device=dev, dtype=torch.uint8)
                s_ = (torch.randn(K // 128, N, generator=gw, device=dev).abs() * 0.01 + 1e-3).to(dtype)
                zs = ((8 - torch.randint(0, 16, (K // 128, N), generator=gw, device=dev)).to(dtype) * s_).to(dtype)
                w = ops.pack_w4(qp, s_, zs)
            elif fmt == "int8g":
                q8 = torch.randint(-128, 128, (K, N), generator=gw, device=dev, dtype=torch.int8)
                s_ = (torch.randn(K // 128, N, generator=gw, device=dev).abs() * 6e-4 + 6e-5).to(dtype)
                zs = ((128 - torch.randint(0, 256, (K // 128, N), generator=gw, device=dev)).to(dtype) * s_).to(dtype)
                w = ops.pack_w8g(q8, s_, zs)
            elif fmt == "int8":
                q8 = torch.randint(-128, 128, (K, N), generator=gw, device=dev, dtype=torch.int8)
                w = ops.pack_w8(q8, (torch.randn(N, generator=gw, device=dev).abs() * 2e-4 + 3e-4).to(dtype))
            else:
                w = ops.pack_f16((torch.randn(K, N, generator=gw, device=dev) * 0.05).to(dtype))
            ws = ops.gemm_workspace(B, [(K, N)], dev)
            gamma = (1 + 0.1 * torch.randn(N, generator=torch.Generator(device=dev).manual_seed(7), device=dev)).to(dtype)
            resid = torch.randn(B, N, generator=torch.Generator(device=dev).manual_seed(9), device=dev).to(dtype)
            r1, r2 = resid.clone(), resid.clone()
            y1, y2 = torch.empty(B, N, device=dev, dtype=dtype), torch.empty(B, N, device=dev, dtype=dtype)
            p1 = ops.wo_gemm(x, w, ws)
            assert comm.all_reduce_norm(p1, r1, gamma, 1e-5, y1)
            p2 = torch.zeros(B, N, device=dev, dtype=dtype)
            for pdl in (False, True):
                r2.copy_(resid)
                comm.gemm_rs(x, w, ws, p2, pdl=pdl)
                comm.gather_norm(p2, r2, gamma, 1e-5, y2)
                torch.cuda.synchronize()
                report(f"gemm_rs + gather_norm == gemm + all_reduce_norm {fmt} {dtype} B{B} K{K} N{N} pdl={int(pdl)}",
                       torch.equal(r1, r2) and torch.equal(y1, y2),
                       f"max |dres| {(r1.float() - r2.float()).abs().max().item():.3g} max |dy| {(y1.float() - y2.float()).abs().max().item():.3g}")
    # the fused pair inside a CUDA graph, three exchanges per replay (odd count: slot parity alternates between replays)
    B, K, N = 32, 512, 4096
    gw = torch.Generator(device=dev).manual_seed(77 + rank)
    x = (torch.randn(B, K, generator=gw, device=dev) * 0.5).half()
    qp = torch.randint(0, 256, (K, N // 2), generator=gw, device=dev, dtype=torch.uint8)
    s_ = (torch.randn(K // 128, N, generator=gw, device=dev).abs() * 0.01 + 1e-3).half()
    zs = ((8 - torch.randint(0, 16, (K // 128, N), generator=gw, device=dev)).half() * s_).half()
    w = ops.pack_w4(qp, s_, zs)
    ws = ops.gemm_workspace(B, [(K, N)], dev)
    gamma = torch.ones(N, device=dev).half()
    resid0 = torch.randn(B, N, generator=torch.Generator(device=dev).manual_seed(3), device=dev).half()
    r_ref, y_ref = resid0.clone(), torch.empty(B, N, device=dev).half()
    for _ in range(3):
        assert comm.all_reduce_norm(ops.wo_gemm(x, w, ws), r_ref, gamma, 1e-5, y_ref)
    torch.cuda.synchronize()
    proj, r_g, y_g = torch.zeros(B, N, device=dev).half(), resid0.clone(), torch.empty(B, N, device=dev).half()

    def chain():
        for _ in range(3):
            comm.gemm_rs(x, w, ws, proj, pdl=True)
            comm.gather_norm(proj, r_g, gamma, 1e-5, y_g)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        chain()
    good = True
    for it in range(5):
        r_g.copy_(resid0)
        graph.replay()
        torch.cuda.synchronize()
        good &= torch.equal(r_g, r_ref) and torch.equal(y_g, y_ref)
    report("graph with 3 x (gemm_rs + gather_norm) replayed 5x", good)

    # vocab-parallel argmax incl. ties across ranks and padded columns
    rows, vloc, vtot = 7, 1008, world * 1008 - 5
    lg = torch.randn(rows, vloc, generator=torch.Generator(device=dev).manual_seed(50 + rank), device=dev).half()
    lg[1, 3] = 50.0                         # the same maximum on every rank: the lowest global index (rank 0) must win
    if rank == world - 1:
        lg[2, vloc - 1] = 99.0              # a padded column (>= vocab_total) must never win
        lg[3, vloc - 6] = 77.0              # last real column
    out = torch.empty(rows, dtype=torch.int32, device=dev)
    comm.argmax(lg, vtot, out)
    parts = [torch.empty_like(lg) for _ in range(world)]
    dist.all_gather(parts, lg)
    full = torch.cat(parts, dim=1)[:, :vtot].float()
    torch.cuda.synchronize()
    report("peer_argmax", torch.equal(out.long(), full.argmax(dim=-1)) and out[1].item() == 3 and out[3].item() == vtot - 1)
    # a graph with an ODD number of all-reduces, replayed: parity must come from the device counter
    t = torch.ones(8, 1024, device=dev).half() * (rank + 1)
    bufs = [t.clone() for _ in range(3)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for b in bufs:
            comm.all_reduce(b)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for b in bufs:
        b.copy_(t)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for b in bufs:
            comm.all_reduce(b)
    good = True
    for it in range(5):
        for b in bufs:
            b.copy_(t)
        graph.replay()
        torch.cuda.synchronize()
        good &= all(bool((b == world * (world + 1) / 2).all()) for b in bufs)
    report("graph with 3 all-reduces replayed 5x", good)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
