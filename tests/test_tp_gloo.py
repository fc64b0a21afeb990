"""CPU test of the N>1 path: world_size 2 over gloo. Each rank evaluates ITS shard of one decoder layer's linear stack
with the oracle (qkv column-parallel by head, o row-parallel, w13 column-parallel, w2 row-parallel -- rtp_llm_b200/tp.py),
all-reduces the row-parallel outputs, and the result must equal the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from rtp_llm_b200 import tp as tpmod

H, HQ, HKV, D, INTER, B = 256, 4, 2, 64, 256, 3


def _weights(fmt):
    g = torch.Generator().manual_seed(11)

    def mk(K, N):
        if fmt == "int4":
            qp = torch.randint(0, 256, (K, N // 2), generator=g, dtype=torch.uint8)
            s = (torch.randn(K // 128, N, generator=g).abs() * 0.01 + 1e-3).half()
            z = torch.randint(0, 16, (K // 128, N), generator=g)
            return ("int4", qp, s, ((8 - z).half() * s).half())
        if fmt == "int8g":
            s = (torch.randn(K // 128, N, generator=g).abs() * 6e-4 + 6e-5).half()
            z = torch.randint(0, 256, (K // 128, N), generator=g)
            return ("int8g", torch.randint(-128, 128, (K, N), generator=g, dtype=torch.int8), s, ((128 - z).half() * s).half())
        if fmt == "int8":
            return ("int8", torch.randint(-128, 128, (K, N), generator=g, dtype=torch.int8),
                    (torch.randn(N, generator=g).abs() * 2e-3 + 1e-4).half(), None)
        return ("f16", (torch.randn(K, N, generator=g) * 0.05).half(), None, None)
    return dict(qkv=mk(H, (HQ + 2 * HKV) * D), o=mk(HQ * D, H), w13=mk(H, 2 * INTER), w2=mk(INTER, H))


def _bits(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def _gemm(x_bits, wt):
    fmt, w, s, zs = wt
    if fmt == "int4":
        return orc.dequant_gemm(x_bits, "int4", w.numpy(), scales=_bits(s), zeros_x_scales=_bits(zs), group=128)
    if fmt == "int8g":
        return orc.dequant_gemm(x_bits, "int8g", w.numpy(), scales=_bits(s), zeros_x_scales=_bits(zs), group=128)
    if fmt == "int8":
        return orc.dequant_gemm(x_bits, "int8", w.numpy(), scales=_bits(s))
    return orc.dequant_gemm(x_bits, "f16", _bits(w))


def _layer(x_bits, W, hq, hkv, inter):
    """qkv -> (take the q slice as a stand-in for attention: linear in heads) -> o ; w13 -> silu*mul -> w2."""
    qkv = _gemm(x_bits, W["qkv"])
    attn = qkv[:, : hq * D]
    proj = orc.from_bits(_gemm(np.ascontiguousarray(attn), W["o"]), False)
    gu = _gemm(x_bits, W["w13"])
    act = orc.silu_and_mul(gu)
    down = orc.from_bits(_gemm(act, W["w2"]), False)
    kv = qkv[:, hq * D:]
    return proj, down, orc.from_bits(np.ascontiguousarray(kv), False)


def _worker(rank, world, port, fmt, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = _weights(fmt)
    x = torch.randn(B, H, generator=torch.Generator().manual_seed(5)).half()
    local = dict(qkv=tpmod.shard_qkv(W["qkv"], HQ, HKV, D, rank, world), o=tpmod.shard_o(W["o"], HQ, D, rank, world),
                 w13=tpmod.shard_w13(W["w13"], INTER, rank, world), w2=tpmod.shard_w2(W["w2"], INTER, rank, world))
    proj, down, kv = _layer(_bits(x), local, HQ // world, max(HKV // world, 1), INTER // world)
    tp = torch.from_numpy(proj.copy())
    td = torch.from_numpy(down.copy())
    dist.all_reduce(tp)
    dist.all_reduce(td)
    kvs = [torch.zeros_like(torch.from_numpy(kv.copy())) for _ in range(world)]
    dist.all_gather(kvs, torch.from_numpy(kv.copy()))
    if rank == 0:
        q.put((tp.numpy(), td.numpy(), [k.numpy() for k in kvs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fmt", ["int4", "int8", "int8g", "f16"])
def test_tp2_sharded_layer_equals_unsharded(fmt):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fmt, q)) for r in range(2)]
    for p in procs:
        p.start()
    proj, down, kvs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W = _weights(fmt)
    x = torch.randn(B, H, generator=torch.Generator().manual_seed(5)).half()
    proj1, down1, kv1 = _layer(_bits(x), W, HQ, HKV, INTER)
    # row-parallel partial sums are rounded to fp16 per rank before the all-reduce (as in the reference, which
    # all-reduces fp16/bf16 tensors): tolerance = a few fp16 ulps of the output magnitude
    np.testing.assert_allclose(proj, proj1, rtol=5e-3, atol=5e-3 * np.abs(proj1).max())
    np.testing.assert_allclose(down, down1, rtol=5e-3, atol=5e-3 * np.abs(down1).max())
    # column-parallel outputs are exact slices: K heads of rank r, V heads of rank r
    hkv = HKV // 2
    k_full, v_full = kv1[:, : HKV * D], kv1[:, HKV * D:]
    for r in range(2):
        np.testing.assert_array_equal(kvs[r][:, : hkv * D], k_full[:, r * hkv * D:(r + 1) * hkv * D])
        np.testing.assert_array_equal(kvs[r][:, hkv * D:], v_full[:, r * hkv * D:(r + 1) * hkv * D])
