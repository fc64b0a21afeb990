#!/usr/bin/env python
"""GPU parity cases (test infrastructure: imported by tests/test_gpu_parity.py, also runnable by hand under gpurun): exercises every kernel against the oracle / the naive GPU
checkers and prints a compact diagnosis per case instead of stopping at the first failure."""
import math
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from rtp_llm_b200 import ops  # noqa: E402
from rtp_llm_b200._lib import B200_FMT_F16, B200_FMT_INT4, B200_FMT_INT8, B200_FMT_INT8G  # noqa: E402
from tests import ref_kernels as refk  # noqa: E402  (naive CUDA-core second opinion, test-only library)

dev = torch.device("cuda:0")
RESULTS = []


def report(name, ok, info=""):
    RESULTS.append((name, ok))
    print(f"[{'PASS' if ok else 'FAIL'}] {name} {info}", flush=True)


def diff_info(got, exp, rtol, atol):
    got = got.float().cpu().numpy().ravel()
    exp = np.asarray(exp, np.float32).ravel()
    err = np.abs(got - exp)
    bad = err > (atol + rtol * np.abs(exp))
    nbad = int(bad.sum()) + int(np.isnan(got).sum())
    info = f"max_abs_err={np.nanmax(err):.4g} bad={nbad}/{got.size}"
    if nbad:
        idx = np.flatnonzero(bad | np.isnan(got))[:6]
        info += " first_bad=" + ", ".join(f"{i}:{got[i]:.4g}/{exp[i]:.4g}" for i in idx)
    return nbad == 0, info


def guard(name, fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        report(name, False, f"EXC {type(e).__name__}: {e}")
        traceback.print_exc()
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------- attention
def make_attn_case(B, Hq, Hkv, T, lens, dtype, seed=0, shuffle=True, D=128):
    g = torch.Generator().manual_seed(seed)
    M = max(math.ceil(L / T) for L in lens)
    npages = sum(math.ceil(L / T) for L in lens) + 1
    pool = torch.randn(npages, 2, Hkv, T, D, generator=g).to(dtype)
    q = torch.randn(B, Hq, D, generator=g).to(dtype)
    perm = (torch.randperm(npages - 1, generator=g) + 1) if shuffle else torch.arange(1, npages)
    block_ids = torch.zeros(B, M, dtype=torch.int32)
    it = iter(perm.tolist())
    for b, L in enumerate(lens):
        for j in range(math.ceil(L / T)):
            block_ids[b, j] = next(it)
    seq = torch.tensor([L - 1 for L in lens], dtype=torch.int32)
    return q, pool, block_ids, seq


def run_attn(q, pool, block_ids, seq, max_len=None):
    qd, poold, bd, sd = q.to(dev), pool.to(dev), block_ids.to(dev), seq.to(dev)
    pl = ops.convert_block_table(bd)
    B, Hq = q.shape[0], q.shape[1]
    max_len = max_len or int(seq.max().item()) + 1
    ws = ops.attn_workspace(B, Hq, pool.shape[2], max_len, dev)
    out = ops.paged_decode_attn(qd, poold, pl, sd, max_len, ws)
    ref = refk.ref_paged_decode_attn(qd, poold, pl, sd)
    torch.cuda.synchronize()
    return out, ref, ws


def attn_vs_oracle(name, B, Hq, Hkv, T, lens, dtype=torch.float16, env=None, D=128):
    def fn():
        old = {}
        for k, v in (env or {}).items():
            old[k] = os.environ.get(k)
            os.environ[k] = str(v)
        try:
            q, pool, block_ids, seq = make_attn_case(B, Hq, Hkv, T, lens, dtype, D=D)
            out, ref, ws = run_attn(q, pool, block_ids, seq)
            is_bf16 = dtype == torch.bfloat16
            bits = lambda t: t.view(torch.int16).numpy().view(np.uint16)
            exp = orc.from_bits(orc.paged_decode_attn(bits(q), bits(pool), orc.convert_block_table(block_ids.numpy()),
                                                      seq.numpy(), Hq, Hkv, D, T, is_bf16=is_bf16), is_bf16)
            tol = 2e-2 if is_bf16 else 1e-2
            ok1, i1 = diff_info(ref, exp, tol, tol)
            report(name + " [ref-kernel vs oracle]", ok1, i1)
            ok2, i2 = diff_info(out, exp, tol, tol)
            report(name + " [kernel vs oracle]", ok2, i2)
            sem_clean = int(ws[: B * Hkv * 4].view(torch.int32).abs().sum().item()) == 0
            report(name + " [semaphores reset]", sem_clean)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    guard(name, fn)


def attn_vs_ref_big(name, B, Hq, Hkv, T, S, dtype=torch.float16, ragged=False):
    """BASELINE-size shapes: the CPU oracle is the checker (it finishes these in well under a second per case with
    OpenMP); the naive GPU kernel is kept as a second opinion."""
    def fn():
        rng = np.random.default_rng(4)
        lens = [int(x) for x in (rng.integers(S // 2, S + 1, B) if ragged else [S] * B)]
        q, pool, block_ids, seq = make_attn_case(B, Hq, Hkv, T, lens, dtype, seed=42)
        out, ref, _ = run_attn(q, pool, block_ids, seq, max_len=S)
        is_bf16 = dtype == torch.bfloat16
        bits = lambda t: t.view(torch.int16).numpy().view(np.uint16)
        exp = orc.from_bits(orc.paged_decode_attn(bits(q), bits(pool), orc.convert_block_table(block_ids.numpy()),
                                                  seq.numpy(), Hq, Hkv, 128, T, is_bf16=is_bf16), is_bf16)
        ok, info = diff_info(out, exp, 1e-2, 1e-2)
        report(name + " [kernel vs oracle]", ok, info)
        ok, info = diff_info(out, ref.float().cpu().numpy(), 1e-2, 1e-2)
        report(name + " [kernel vs ref-kernel]", ok, info)
    guard(name, fn)


# ---------------------------------------------------------------------------------------------- GEMM
def make_w4(K, N, seed=1, simple=False):
    rng = np.random.default_rng(seed)
    qp = rng.integers(0, 256, (K, N // 2)).astype(np.uint8)
    G = K // 128
    if simple:
        s = np.ones((G, N), np.float16)
        zs = np.zeros((G, N), np.float16)
    else:
        s = (np.abs(rng.standard_normal((G, N))) * 0.01 + 1e-3).astype(np.float16)
        z = rng.integers(0, 16, (G, N))
        zs = ((8 - z).astype(np.float16) * s).astype(np.float16)
    return qp, s, zs


def make_w8g(K, N, seed=1, simple=False):
    rng = np.random.default_rng(seed)
    q = rng.integers(-128, 128, (K, N)).astype(np.int8)
    G = K // 128
    if simple:
        s, zs = np.ones((G, N), np.float16), np.zeros((G, N), np.float16)
    else:   # realistic magnitude: W' ~ 0.02 like an 8-bit GPTQ layer (range / 255 per group)
        s = (np.abs(rng.standard_normal((G, N))) * 2e-4 + 3e-4).astype(np.float16)
        z = rng.integers(0, 256, (G, N))
        zs = ((128 - z).astype(np.float32) * s.astype(np.float32)).astype(np.float16)
    return q, s, zs


def gemm_case(name, fmt, B, K, N, dtype=torch.float16, simple=False, onehot=False, env=None, bias=False, big=False):
    def fn():
        old = {}
        for k, v in (env or {}).items():
            old[k] = os.environ.get(k)
            os.environ[k] = str(v)
        try:
            rng = np.random.default_rng(2)
            if onehot:
                xf = np.zeros((B, K), np.float32)
                for b in range(B):
                    xf[b, (b * 37 + 5) % K] = 1.0
            else:
                xf = rng.standard_normal((B, K)).astype(np.float32)
            x = torch.from_numpy(xf).to(dtype).to(dev)
            bias_t = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(dtype).to(dev) if bias else None
            if fmt == B200_FMT_INT4:
                qp, s, zs = make_w4(K, N, simple=simple)
                qd = torch.from_numpy(qp).to(dev)
                sd, zd = torch.from_numpy(s).to(dtype).to(dev), torch.from_numpy(zs).to(dtype).to(dev)
                w = ops.pack_w4(qd, sd, zd)
                ref = refk.ref_dequant_gemm(x, fmt, qd, sd, zd, 128, bias_t)
            elif fmt == B200_FMT_INT8G:
                q8, s, zs = make_w8g(K, N, simple=simple)
                qd = torch.from_numpy(q8).to(dev)
                sd, zd = torch.from_numpy(s).to(dtype).to(dev), torch.from_numpy(zs).to(dtype).to(dev)
                w = ops.pack_w8g(qd, sd, zd)
                ref = refk.ref_dequant_gemm(x, fmt, qd, sd, zd, 128, bias_t)
            elif fmt == B200_FMT_INT8:
                q8 = rng.integers(-128, 128, (K, N)).astype(np.int8)
                # realistic magnitude: W' ~ N(0, 0.02) like the loader's scale = amax/128 (device_impl.py:190)
                s = (np.abs(rng.standard_normal(N)) * 2e-4 + 3e-4).astype(np.float32)
                if simple:
                    s[:] = 1.0
                qd, sd = torch.from_numpy(q8).to(dev), torch.from_numpy(s).to(dtype).to(dev)
                w = ops.pack_w8(qd, sd)
                ref = refk.ref_dequant_gemm(x, fmt, qd, sd, None, 128, bias_t)
            else:
                wf = (rng.standard_normal((K, N)) * 0.05).astype(np.float32)
                wd = torch.from_numpy(wf).to(dtype).to(dev)
                w = ops.pack_f16(wd)
                ref = refk.ref_dequant_gemm(x, fmt, wd, None, None, 128, bias_t)
            ws = ops.gemm_workspace(max(B, 1), [(K, N)], dev)
            y = ops.wo_gemm(x, w, ws, bias=bias_t)
            torch.cuda.synchronize()
            tol = 2e-2 if fmt == B200_FMT_INT4 else 1e-2
            if dtype == torch.bfloat16:
                tol = 4e-2
            ok, info = diff_info(y, ref.float().cpu().numpy(), tol, tol * (1 if not simple else 1))
            report(name + " [kernel vs ref-kernel]", ok, info)
            if not onehot:     # every size, BASELINE shapes included: the CPU oracle is the checker
                bits = lambda t: t.cpu().view(torch.int16).numpy().view(np.uint16)
                is_bf16 = dtype == torch.bfloat16
                if fmt == B200_FMT_INT4:
                    exp = orc.dequant_gemm(bits(x), "int4", qp, scales=bits(sd), zeros_x_scales=bits(zd), group=128,
                                           bias=bits(bias_t) if bias else None, is_bf16=is_bf16)
                elif fmt == B200_FMT_INT8G:
                    exp = orc.dequant_gemm(bits(x), "int8g", q8, scales=bits(sd), zeros_x_scales=bits(zd), group=128,
                                           bias=bits(bias_t) if bias else None, is_bf16=is_bf16)
                elif fmt == B200_FMT_INT8:
                    exp = orc.dequant_gemm(bits(x), "int8", q8, scales=bits(sd), bias=bits(bias_t) if bias else None,
                                           is_bf16=is_bf16)
                else:
                    exp = orc.dequant_gemm(bits(x), "f16", bits(wd), bias=bits(bias_t) if bias else None, is_bf16=is_bf16)
                ok2, i2 = diff_info(y, orc.from_bits(exp, is_bf16), tol, tol)
                report(name + " [kernel vs oracle]", ok2, i2)
            sem_clean = int(ws[:16384].view(torch.int32).abs().sum().item()) == 0
            report(name + " [semaphores reset]", sem_clean)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    guard(name, fn)


def gemm_silu_case(name, fmt, B, K, inter, dtype=torch.float16, env=None):
    """w13 GEMM with the fused SiLU*mul epilogue vs oracle (GEMM on the un-permuted weight, then silu_and_mul)."""
    def fn():
        old = {}
        for k, v in (env or {}).items():
            old[k] = os.environ.get(k)
            os.environ[k] = str(v)
        try:
            rng = np.random.default_rng(9)
            N = 2 * inter
            x = torch.from_numpy(rng.standard_normal((B, K)).astype(np.float32)).to(dtype).to(dev)
            bits = lambda t: t.cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
            if fmt == B200_FMT_INT4:
                qp, s, zs = make_w4(K, N)
                qd, sd, zd = torch.from_numpy(qp).to(dev), torch.from_numpy(s).to(dtype).to(dev), torch.from_numpy(zs).to(dtype).to(dev)
                w = ops.pack_w4(ops.interleave_gate_up(qd, inter, packed_int4=True), ops.interleave_gate_up(sd, inter),
                                ops.interleave_gate_up(zd, inter))
                gu = orc.dequant_gemm(bits(x), "int4", qp, scales=bits(sd), zeros_x_scales=bits(zd), group=128)
            elif fmt == B200_FMT_INT8G:
                q8, s, zs = make_w8g(K, N)
                qd, sd, zd = torch.from_numpy(q8).to(dev), torch.from_numpy(s).to(dtype).to(dev), torch.from_numpy(zs).to(dtype).to(dev)
                w = ops.pack_w8g(ops.interleave_gate_up(qd, inter), ops.interleave_gate_up(sd, inter), ops.interleave_gate_up(zd, inter))
                gu = orc.dequant_gemm(bits(x), "int8g", q8, scales=bits(sd), zeros_x_scales=bits(zd), group=128)
            elif fmt == B200_FMT_INT8:
                q8 = rng.integers(-128, 128, (K, N)).astype(np.int8)
                s = (np.abs(rng.standard_normal(N)) * 2e-4 + 3e-4).astype(np.float32)
                qd, sd = torch.from_numpy(q8).to(dev), torch.from_numpy(s).to(dtype).to(dev)
                w = ops.pack_w8(ops.interleave_gate_up(qd, inter), ops.interleave_gate_up(sd, inter))
                gu = orc.dequant_gemm(bits(x), "int8", q8, scales=bits(sd))
            else:
                wd = torch.from_numpy((rng.standard_normal((K, N)) * 0.05).astype(np.float32)).to(dtype).to(dev)
                w = ops.pack_f16(ops.interleave_gate_up(wd, inter))
                gu = orc.dequant_gemm(bits(x), "f16", bits(wd))
            exp = orc.from_bits(orc.silu_and_mul(gu), False)
            ws = ops.gemm_workspace(B, [(K, N)], dev)
            y = ops.wo_gemm(x, w, ws, silu_mul=True)
            torch.cuda.synchronize()
            ok, info = diff_info(y, exp, 2e-2, 2e-2)
            report(name + " [fused silu*mul vs oracle]", ok and tuple(y.shape) == (B, inter), info)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    guard(name, fn)


def glue_cases():
    def fn():
        rng = np.random.default_rng(3)
        for dtype in (torch.float16, torch.bfloat16):
            is_bf16 = dtype == torch.bfloat16
            bits = lambda t: t.cpu().view(torch.int16).numpy().view(np.uint16)
            x = torch.from_numpy(rng.standard_normal((5, 4096)).astype(np.float32)).to(dtype).to(dev)
            r = torch.from_numpy(rng.standard_normal((5, 4096)).astype(np.float32)).to(dtype).to(dev)
            gmm = torch.from_numpy(rng.standard_normal(4096).astype(np.float32)).to(dtype).to(dev)
            r2 = r.clone()
            y = ops.add_rmsnorm(x, r2, gmm, 1e-6)
            ye, re_ = orc.add_rmsnorm(bits(x), bits(r), bits(gmm), 1e-6, is_bf16)
            ok, info = diff_info(y, orc.from_bits(ye, is_bf16), 2e-2, 2e-2)
            report(f"add_rmsnorm {dtype}", ok and bool((bits(r2) == re_).all()), info)
            gu = torch.from_numpy(rng.standard_normal((5, 2 * 1024)).astype(np.float32)).to(dtype).to(dev)
            s = ops.silu_and_mul(gu)
            ok, info = diff_info(s, orc.from_bits(orc.silu_and_mul(bits(gu), is_bf16), is_bf16), 2e-2, 2e-2)
            report(f"silu_and_mul {dtype}", ok, info)
            B, Hq, Hkv, D, T, M = 3, 8, 2, 128, 16, 4
            qkv = torch.from_numpy(rng.standard_normal((B, (Hq + 2 * Hkv) * D)).astype(np.float32)).to(dtype).to(dev)
            pool = torch.zeros(1 + B * M, 2, Hkv, T, D, dtype=dtype, device=dev)
            block_ids = (torch.arange(B * M, dtype=torch.int32) + 1).reshape(B, M).to(dev)
            pl = ops.convert_block_table(block_ids)
            seq = torch.tensor([0, 37, 63], dtype=torch.int32, device=dev)
            qo = ops.rope_append(qkv, pool, pl, seq, Hq, 500000.0)
            qe, poole = orc.rope_append(bits(qkv), bits(torch.zeros_like(pool)), pl.cpu().numpy(), seq.cpu().numpy(), Hq,
                                        Hkv, D, T, 500000.0, is_bf16)
            ok1, i1 = diff_info(qo, orc.from_bits(qe, is_bf16), 2e-2, 2e-2)
            ok2, i2 = diff_info(pool, orc.from_bits(poole, is_bf16), 2e-2, 2e-2)
            report(f"rope_append {dtype}", ok1 and ok2, i1 + " | " + i2)
        lg = torch.randn(7, 128256, device=dev)
        lg[2, 10] = lg[2, 500] = 99.0
        am = ops.argmax(lg)
        report("argmax fp32", bool((am.cpu().numpy() == orc.argmax(lg.cpu().numpy())).all()))
        for dt in (torch.float16, torch.bfloat16):
            lh = lg.to(dt)                                   # rounding creates many exact ties: lowest index must win
            ah = ops.argmax(lh)
            report(f"argmax {dt} (vectorised path, ties)", bool((ah.cpu().numpy() == orc.argmax(lh.float().cpu().numpy())).all()))
            lo = lh[:, :1001].contiguous()                   # vocab % 8 != 0 -> scalar path
            report(f"argmax {dt} (scalar path)", bool((ops.argmax(lo).cpu().numpy() == orc.argmax(lo.float().cpu().numpy())).all()))
        tbl = torch.randn(1000, 4096, device=dev).half()
        ids = torch.tensor([3, 999, 0], dtype=torch.int32, device=dev)
        report("embedding", bool((ops.embedding(ids, tbl) == tbl[ids.long()]).all()))
        bid = torch.randint(0, 512, (4, 8), dtype=torch.int32)
        plan = ops.paged_attn_plan(torch.tensor([9, 19, 64, 129], dtype=torch.int32, device=dev),
                                   torch.arange(4 * 9, dtype=torch.int32).reshape(4, 9).to(dev), 16)
        pe = orc.paged_attn_plan(np.array([9, 19, 64, 129], np.int32), np.arange(36, dtype=np.int32).reshape(4, 9), 16)
        ok = all(np.array_equal(plan[k].cpu().numpy()[: len(pe[k])], pe[k]) for k in pe)
        report("paged_attn_plan", ok)
        report("convert_block_table", np.array_equal(ops.convert_block_table(bid.to(dev)).cpu().numpy(),
                                                     orc.convert_block_table(bid.numpy())))
    guard("glue", fn)


if __name__ == "__main__":
    which = sys.argv[1:] or ["attn", "gemm", "glue"]
    ops.device_check(0)
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    if "glue" in which:
        glue_cases()
    if "attn" in which:
        attn_vs_oracle("attn B1 len1 p64 g4", 1, 4, 1, 64, [1])
        attn_vs_oracle("attn B2 p64 g4 [10,64]", 2, 8, 2, 64, [10, 64])
        attn_vs_oracle("attn B4 p64 g4 [65,128,200,513]", 4, 8, 2, 64, [65, 128, 200, 513])
        attn_vs_oracle("attn p16 g4", 4, 8, 2, 16, [10, 20, 65, 130])
        attn_vs_oracle("attn p32 mha", 3, 4, 4, 32, [64, 65, 1])
        attn_vs_oracle("attn p128 g8", 3, 16, 2, 128, [127, 129, 300])
        attn_vs_oracle("attn p64 g16", 2, 16, 1, 64, [100, 257])
        attn_vs_oracle("attn p64 g4 bf16", 3, 8, 2, 64, [63, 64, 300], dtype=torch.bfloat16)
        attn_vs_oracle("attn forced split 1 tile", 3, 8, 2, 64, [63, 64, 700], env={"B200_ATTN_TILES_PER_SPLIT": 1})
        attn_vs_oracle("attn forced nosplit", 3, 8, 2, 64, [63, 64, 700], env={"B200_ATTN_TILES_PER_SPLIT": 1000})
        attn_vs_ref_big("attn Llama B32 S2048", 32, 32, 8, 64, 2048)
        attn_vs_ref_big("attn Llama B32 S2048 ragged", 32, 32, 8, 64, 2048, ragged=True)
        attn_vs_ref_big("attn Qwen72B-TP8 B16 S8192", 16, 8, 1, 64, 8192)
    fmts = [(f, n) for f, n in ((B200_FMT_F16, "f16"), (B200_FMT_INT8, "int8"), (B200_FMT_INT4, "int4"))
            if "gemm" in which or f"gemm:{n}" in which]
    if fmts:
        for fmt, nm in fmts:
            gemm_case(f"gemm {nm} onehot simple B16 K128 N128", fmt, 16, 128, 128, simple=True, onehot=True)
            gemm_case(f"gemm {nm} simple B16 K128 N128", fmt, 16, 128, 128, simple=True)
            gemm_case(f"gemm {nm} B16 K256 N256", fmt, 16, 256, 256)
            gemm_case(f"gemm {nm} B5 K512 N384 bias", fmt, 5, 512, 384, bias=True)
            gemm_case(f"gemm {nm} B32 K1024 N256 splitk4", fmt, 32, 1024, 256, env={"B200_GEMM_SPLITK": 4})
            gemm_case(f"gemm {nm} B19 K1024 N256 splitk8 semaphore", fmt, 19, 1024, 256,
                      env={"B200_GEMM_SPLITK": 8, "B200_GEMM_CLUSTER": 0})
            gemm_case(f"gemm {nm} B19 K1024 N384 splitk3 bias cluster", fmt, 19, 1024, 384, env={"B200_GEMM_SPLITK": 3}, bias=True)
            gemm_case(f"gemm {nm} B33 K512 N200 (ragged N, bpad64)", fmt, 33, 512, 200)
            gemm_case(f"gemm {nm} B100 K512 N256 (bpad128)", fmt, 100, 512, 256)
            gemm_case(f"gemm {nm} bf16 B8 K256 N256", fmt, 8, 256, 256, dtype=torch.bfloat16)
            gemm_silu_case(f"gemm {nm} silu B7 K256 I192", fmt, 7, 256, 192)
            gemm_silu_case(f"gemm {nm} silu B32 K1024 I128 cluster split 4", fmt, 32, 1024, 128, env={"B200_GEMM_SPLITK": 4})
    if "gemm" in which or "gemm:big" in which:
        gemm_case("gemm int4 Llama qkv B32", B200_FMT_INT4, 32, 4096, 6144, big=True)
        gemm_case("gemm int4 Llama w2 B32", B200_FMT_INT4, 32, 14336, 4096, big=True)
        gemm_case("gemm int4 Llama w13 B64", B200_FMT_INT4, 64, 4096, 28672, big=True)
        gemm_case("gemm f16 lm_head-ish B32", B200_FMT_F16, 32, 4096, 16032, big=True)
    nfail = sum(1 for _, ok in RESULTS if not ok)
    print(f"SUMMARY: {len(RESULTS) - nfail} passed, {nfail} failed, {time.time() - t0:.1f}s", flush=True)
    sys.exit(1 if nfail else 0)
