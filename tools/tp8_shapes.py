#!/usr/bin/env python
"""Per-kernel timings at the PER-RANK shapes of Llama-3-8B TP=8 / TP=2 (developer tool, one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kernel_bench import bench_attn, bench_gemm
from rtp_llm_b200._lib import B200_FMT_INT4, B200_FMT_F16
print("--- TP8 per-rank shapes")
bench_attn(32, 4, 1, 2048)
for c in (1, 2, 4, 8, 16, 32):
    bench_attn(32, 4, 1, 2048, env={"B200_ATTN_TILES_PER_SPLIT": c})
bench_gemm(B200_FMT_INT4, 32, 4096, 768)
bench_gemm(B200_FMT_INT4, 32, 512, 4096)
bench_gemm(B200_FMT_INT4, 32, 4096, 3584)
bench_gemm(B200_FMT_INT4, 32, 1792, 4096)
bench_gemm(B200_FMT_F16, 32, 4096, 16032)
for s in (1, 2, 4, 8):
    bench_gemm(B200_FMT_INT4, 32, 4096, 768, env={"B200_GEMM_SPLITK": s})
    bench_gemm(B200_FMT_INT4, 32, 4096, 3584, env={"B200_GEMM_SPLITK": s})
    bench_gemm(B200_FMT_INT4, 32, 1792, 4096, env={"B200_GEMM_SPLITK": s})
print("--- TP2 per-rank shapes")
bench_attn(32, 16, 4, 2048)
for c in (8, 16, 32):
    bench_attn(32, 16, 4, 2048, env={"B200_ATTN_TILES_PER_SPLIT": c})
