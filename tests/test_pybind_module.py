"""The pybind plugin form of the boundary (rtp_llm_b200/csrc/pybind_ops.cc): `registerPyModuleOps` fills an `rtp_llm_ops`
submodule with an XQAAttnOp-shaped class. CPU part: it builds, imports and exposes the reference's method set; GPU part:
the op runs through the same C ABI and matches the oracle."""
import importlib.util
import math
import os

import numpy as np
import pytest
import torch

from rtp_llm_b200 import build


def _load():
    path = build.build_pybind()
    spec = importlib.util.spec_from_file_location("b200_compute_ops", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.rtp_llm_ops


def test_pybind_module_exposes_the_reference_surface():
    ops = _load()
    for m in ("support", "prepare", "update", "update_kv_cache_offset", "forward"):      # XQAAttnOp.cc:158-176
        assert hasattr(ops.B200AttnOp, m)
    assert hasattr(ops.B200AttnParams, "kv_cache_offset") and hasattr(ops.B200AttnParams, "__cpp_ptr__")
    cfg = ops.AttentionConfigs()
    cfg.head_num, cfg.kv_head_num, cfg.size_per_head, cfg.tokens_per_block = 32, 8, 128, 64
    inp = ops.PyAttentionInputs()
    inp.is_prefill = True
    assert ops.B200AttnOp(cfg).support(inp) is False       # decode only (and no sm_100 device in the CPU container)
    assert issubclass(ops.B200AttnParams, ops.ParamsBase)


@pytest.mark.gpu
def test_pybind_attn_and_linear_ops_match_oracle():
    from oracle import oracle as orc
    from rtp_llm_b200 import ops as b200ops
    from rtp_llm_b200._lib import B200_FMT_INT4
    ops = _load()
    dev = torch.device("cuda")
    torch.manual_seed(42)
    B, Hq, Hkv, D, T = 4, 8, 2, 128, 64
    lens = [63, 64, 200, 511]
    M = max(math.ceil((L + 1) / T) for L in lens)
    P = B * M + 1
    pool = torch.randn(P, 2, Hkv, T, D, device=dev).half()
    block_ids = (torch.randperm(P - 1, device=dev).to(torch.int32) + 1).reshape(B, M)
    seq = torch.tensor(lens, dtype=torch.int32, device=dev)
    qkv = torch.randn(B, (Hq + 2 * Hkv) * D, device=dev).half()
    cfg = ops.AttentionConfigs()
    cfg.head_num, cfg.kv_head_num, cfg.size_per_head, cfg.tokens_per_block = Hq, Hkv, D, T
    cfg.max_seq_len = M * T
    cfg.rope_config.base = 500000.0
    inp = ops.PyAttentionInputs()
    inp.sequence_lengths, inp.kv_cache_kernel_block_id_device = seq, block_ids
    op = ops.B200AttnOp(cfg)
    assert op.support(inp)
    params = op.prepare(inp)
    kv = ops.LayerKVCache()
    kv.kv_cache_base = pool
    pool0 = pool.clone()
    q = ops.b200_rope_kvcache_decode(qkv, kv, params, cfg)
    out = op.forward(q, kv, params)
    torch.cuda.synchronize()
    bits = lambda t: t.cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
    pl = orc.convert_block_table(block_ids.cpu().numpy())
    assert np.array_equal(params.kv_cache_offset.cpu().numpy(), pl)
    q_e, pool_e = orc.rope_append(bits(qkv), bits(pool0), pl, np.array(lens, np.int32), Hq, Hkv, D, T, 500000.0)
    out_e = orc.paged_decode_attn(q_e.reshape(B, Hq, D), pool_e, pl, np.array(lens, np.int32), Hq, Hkv, D, T)
    np.testing.assert_allclose(out.float().cpu().numpy(), orc.from_bits(out_e, False), rtol=1e-2, atol=1e-2)
    # update(): new block table, refreshed in place
    inp2 = ops.PyAttentionInputs()
    inp2.sequence_lengths, inp2.kv_cache_kernel_block_id_device = seq, block_ids.flip(0).contiguous()
    op.update(params, inp2)
    torch.cuda.synchronize()
    assert np.array_equal(params.kv_cache_offset.cpu().numpy(), orc.convert_block_table(inp2.kv_cache_kernel_block_id_device.cpu().numpy()))
    with pytest.raises(RuntimeError):
        op.forward(q, None, params)                         # "decode should have kv cache."
    # linear op on a packed INT4 weight
    rng = np.random.default_rng(1)
    K, N = 256, 256
    qp = rng.integers(0, 256, (K, N // 2)).astype(np.uint8)
    s = (np.abs(rng.standard_normal((K // 128, N))) * 0.01 + 1e-3).astype(np.float16)
    zs = ((8 - rng.integers(0, 16, (K // 128, N))).astype(np.float16) * s).astype(np.float16)
    pw = b200ops.pack_w4(torch.from_numpy(qp).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(zs).to(dev))
    lin = ops.B200LinearOp(B200_FMT_INT4, K, N, pw.data)
    x = torch.randn(5, K, device=dev).half()
    y = lin.forward(x)
    exp = orc.dequant_gemm(bits(x), "int4", qp, scales=s, zeros_x_scales=zs, group=128)
    np.testing.assert_allclose(y.float().cpu().numpy(), orc.from_bits(exp, False), rtol=2e-2, atol=2e-2)
