#!/usr/bin/env python
"""Per-kernel timing at the BASELINE shapes (developer tool, run under gpurun). CUDA events on the launching stream,
warm-up, inputs rotated through > L2-size working sets so nothing is served from the 126 MB L2."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtp_llm_b200 import ops  # noqa: E402
from rtp_llm_b200._lib import B200_FMT_F16, B200_FMT_INT4, B200_FMT_INT8  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6572.5
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def timeit(fn, n_rot, iters=20, warm=3):
    """Average device time per call with the calls replayed from a CUDA graph (no CPU launch overhead in the number)."""
    calls = max(n_rot, 8)
    for i in range(calls):
        fn(i % n_rot)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(calls):
            fn(i % n_rot)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(calls):
            fn(i % n_rot)
    for _ in range(warm):
        g.replay()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        g.replay()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / (iters * calls) * 1e3  # us


def bench_attn(B, Hq, Hkv, S, T=64, nrot=3, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    M = math.ceil(S / T)
    P = B * M + 1
    pools = [torch.randn(P, 2, Hkv, T, 128, device=dev).half() for _ in range(nrot)]
    q = torch.randn(B, Hq, 128, device=dev).half()
    bid = (torch.randperm(P - 1, device=dev).to(torch.int32) + 1).reshape(B, M)
    pl = ops.convert_block_table(bid)
    seq = torch.full((B,), S - 1, dtype=torch.int32, device=dev)
    ws = ops.attn_workspace(B, Hq, Hkv, S, dev)
    out = torch.empty(B, Hq * 128, device=dev).half()
    us = timeit(lambda i: ops.paged_decode_attn(q, pools[i], pl, seq, S, ws, out=out), nrot)
    byt = 2 * B * S * Hkv * 128 * 2
    for k in (env or {}):
        os.environ.pop(k, None)
    print(f"attn B{B} Hq{Hq} Hkv{Hkv} S{S} {env or ''}: {us:8.1f} us  {byt / us / 1e3:7.0f} GB/s  frac={byt / us / 1e3 / PEAK:.3f}", flush=True)
    return us


def bench_gemm(fmt, B, K, N, env=None, pdl=False):
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    nm = {B200_FMT_F16: "f16", B200_FMT_INT8: "int8", B200_FMT_INT4: "int4"}[fmt]
    wbytes = {B200_FMT_F16: 2 * K * N, B200_FMT_INT8: K * N + 2 * N, B200_FMT_INT4: K * N / 2 + K * N / 128 * 4}[fmt]
    nrot = max(2, int(300e6 // wbytes) + 1)
    ws_ = []
    for r in range(nrot):
        if fmt == B200_FMT_INT4:
            qp = torch.randint(0, 256, (K, N // 2), device=dev, dtype=torch.uint8)
            s = (torch.randn(K // 128, N, device=dev).abs() * 0.01 + 1e-3).half()
            w = ops.pack_w4(qp, s, s)
        elif fmt == B200_FMT_INT8:
            w = ops.pack_w8(torch.randint(-128, 128, (K, N), device=dev, dtype=torch.int8), torch.ones(N, device=dev).half())
        else:
            w = ops.PackedWeight(B200_FMT_F16, K, N, (torch.randn(N, K, device=dev) * 0.02).half())
        ws_.append(w)
    x = torch.randn(B, K, device=dev).half()
    y = torch.empty(B, N, device=dev).half()
    wk = ops.gemm_workspace(B, [(K, N)], dev)
    us = timeit(lambda i: ops.wo_gemm(x, ws_[i], wk, out=y, pdl=pdl), nrot)
    byt = wbytes + 2 * B * (K + N)
    for k in (env or {}):
        os.environ.pop(k, None)
    print(f"gemm {nm} B{B} K{K} N{N} {env or ''}{' pdl' if pdl else ''}: {us:8.1f} us  {byt / us / 1e3:7.0f} GB/s  frac={byt / us / 1e3 / PEAK:.3f}",
          flush=True)
    return us


if __name__ == "__main__":
    which = sys.argv[1:] or ["attn", "gemm"]
    if which[0] == "one_gemm":     # one_gemm fmt B K N  (for ncu)
        f = {"f16": B200_FMT_F16, "int8": B200_FMT_INT8, "int4": B200_FMT_INT4}[which[1]]
        bench_gemm(f, int(which[2]), int(which[3]), int(which[4]))
        sys.exit(0)
    if which[0] == "one_attn":     # one_attn B Hq Hkv S
        bench_attn(int(which[1]), int(which[2]), int(which[3]), int(which[4]))
        sys.exit(0)
    print(torch.cuda.get_device_name(0), "peak", PEAK)
    if "attn" in which:
        bench_attn(32, 32, 8, 2048)
        for c in (2, 4, 8, 16, 32):
            bench_attn(32, 32, 8, 2048, env={"B200_ATTN_TILES_PER_SPLIT": c})
        bench_attn(1, 32, 8, 4096)
        bench_attn(64, 32, 8, 4096)
        bench_attn(16, 8, 1, 8192)
    if "gemm" in which:
        for B in (32,):
            bench_gemm(B200_FMT_INT4, B, 4096, 6144)
            bench_gemm(B200_FMT_INT4, B, 4096, 4096)
            bench_gemm(B200_FMT_INT4, B, 4096, 28672)
            bench_gemm(B200_FMT_INT4, B, 14336, 4096)
        bench_gemm(B200_FMT_INT4, 32, 4096, 28672, pdl=True)
        bench_gemm(B200_FMT_INT4, 1, 4096, 28672)
        bench_gemm(B200_FMT_INT4, 64, 4096, 28672)
        bench_gemm(B200_FMT_INT8, 32, 4096, 28672)
        bench_gemm(B200_FMT_F16, 32, 4096, 28672)
        bench_gemm(B200_FMT_F16, 32, 4096, 128256)
