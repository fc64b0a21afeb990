#!/usr/bin/env python
"""Head-to-head on one B200: the reference's Blackwell decode-attention kernel vs b200_paged_decode_attn.

The reference arm is flashinfer.decode.trtllm_batch_decode_with_kv_cache called exactly as
rtp_llm/models_py/modules/factory/attention/cuda_impl/trtllm_gen.py:503-546 does (HND cache [P,2,Hkv,T,D], block table of
page ids, seq_lens = sequence_lengths + 1, bmm1_scale = D^-1/2, bmm2_scale 1, window_left -1, no sinks). Both arms are
timed the same way: CUDA-graph replays over rotating KV pools (working set > L2), CUDA events, after warm-up.
Developer tool (run under gpurun); writes one table to stdout.
"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtp_llm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6572.5
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def timeit(fn, n_rot, iters=20, warm=3):
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


def case(B, Hq, Hkv, S, T=64):
    M = math.ceil(S / T)
    P = B * M + 1
    byt = 2 * B * S * Hkv * 128 * 2
    nrot = max(2, int(400e6 // byt) + 1)
    nrot = min(nrot, 24)
    pools = [torch.randn(P, 2, Hkv, T, 128, device=dev).half() for _ in range(nrot)]
    q = torch.randn(B, Hq, 128, device=dev).half()
    bid = (torch.randperm(P - 1, device=dev).to(torch.int32) + 1).reshape(B, M).contiguous()
    pl = ops.convert_block_table(bid)
    seq = torch.full((B,), S - 1, dtype=torch.int32, device=dev)
    ws = ops.attn_workspace(B, Hq, Hkv, S, dev)
    out = torch.empty(B, Hq * 128, device=dev).half()
    ours = timeit(lambda i: ops.paged_decode_attn(q, pools[i], pl, seq, S, ws, out=out), nrot)
    o_ours = ops.paged_decode_attn(q, pools[0], pl, seq, S, ws).float()

    theirs, err, note = float("nan"), float("nan"), ""
    try:
        import flashinfer
        wsb = torch.zeros(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        seq1 = (seq + 1).contiguous()
        scale = 1.0 / math.sqrt(128)

        def ref(i):
            return flashinfer.decode.trtllm_batch_decode_with_kv_cache(
                query=q, kv_cache=pools[i], workspace_buffer=wsb, block_tables=bid, seq_lens=seq1, max_seq_len=S,
                bmm1_scale=scale, bmm2_scale=1.0, window_left=-1, sinks=None, out_dtype=torch.float16, q_len_per_req=1)
        t0 = time.time()
        o_ref = ref(0)
        torch.cuda.synchronize()
        note = f"first call {time.time() - t0:.0f}s"
        err = (o_ref.reshape(B, -1).float() - o_ours).abs().max().item()
        theirs = timeit(lambda i: ref(i), nrot)
    except Exception as e:  # noqa: BLE001
        note = f"trtllm-gen unavailable: {type(e).__name__}: {str(e)[:200]}"
    print(f"B{B:<3d} Hq{Hq:<3d} Hkv{Hkv:<2d} S{S:<5d} | ours {ours:8.1f} us {byt / ours / 1e3:6.0f} GB/s frac {byt / ours / 1e3 / PEAK:.3f}"
          f" | trtllm-gen {theirs:8.1f} us {byt / theirs / 1e3:6.0f} GB/s frac {byt / theirs / 1e3 / PEAK:.3f}"
          f" | ours/theirs time {ours / theirs:.3f} | max|diff| {err:.2e} | {note}", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "HBM peak", PEAK, "GB/s; fp16 KV, page 64, head_dim 128")
    for (B, Hq, Hkv, S) in [(32, 32, 8, 2048), (1, 32, 8, 4096), (64, 32, 8, 4096), (32, 4, 1, 2048), (16, 8, 1, 8192),
                            (8, 32, 8, 4096)]:
        case(B, Hq, Hkv, S)
