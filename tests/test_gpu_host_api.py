"""GPU tests of the reference-shaped host API (attention.B200DecodeImpl, linear.B200WeightOnlyLinear, device.B200Impl):
the tests read like the reference's own (base_attention_test.py / fp8_linear_test.py): build inputs, run the impl,
compare with a de-quantised / un-paged oracle."""
import math
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402
from rtp_llm_b200 import attention, device, linear  # noqa: E402


def _cfg(Hq, Hkv, T, max_seq_len):
    return types.SimpleNamespace(head_num=Hq, kv_head_num=Hkv, size_per_head=128, tokens_per_block=T,
                                 kernel_tokens_per_block=T, max_seq_len=max_seq_len, need_rope_kv_cache=True,
                                 rope_config=types.SimpleNamespace(base=500000.0), q_scaling=1.0, kv_cache_dtype="BASE")


def test_decode_impl_rope_append_then_attention_and_graph_refresh():
    dev = torch.device("cuda")
    torch.manual_seed(42)                                   # base_attention_test.py:165
    B, Hq, Hkv, D, T = 4, 8, 2, 128, 16
    lens = [64, 128, 256, 512][:B]                          # KV length BEFORE this token (sequence_lengths)
    lens = [9, 19, 64, 129]
    M = max(math.ceil((L + 1) / T) for L in lens)
    P = B * M + 1
    pool = torch.randn(P, 2, Hkv, T, D, device=dev).half()
    block_ids = (torch.randperm(P - 1, device=dev).to(torch.int32) + 1).reshape(B, M)
    seq = torch.tensor(lens, dtype=torch.int32, device=dev)
    qkv = torch.randn(B, (Hq + 2 * Hkv) * D, device=dev).half()
    inputs = types.SimpleNamespace(sequence_lengths=seq, kv_cache_kernel_block_id_device=block_ids, is_prefill=False)
    cfg = _cfg(Hq, Hkv, T, M * T)
    assert attention.B200DecodeImpl.support(cfg, inputs)
    impl = attention.B200DecodeImpl(cfg, inputs)
    assert impl.support_cuda_graph()
    pool0 = pool.clone()
    out = impl.forward(qkv, types.SimpleNamespace(kv_cache_base=pool), 0)
    torch.cuda.synchronize()
    bits = lambda t: t.cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
    pl = orc.convert_block_table(block_ids.cpu().numpy())
    q_e, pool_e = orc.rope_append(bits(qkv), bits(pool0), pl, np.array(lens, np.int32), Hq, Hkv, D, T, 500000.0)
    out_e = orc.paged_decode_attn(q_e.reshape(B, Hq, D), pool_e, pl, np.array(lens, np.int32), Hq, Hkv, D, T)
    np.testing.assert_allclose(out.float().cpu().numpy(), orc.from_bits(out_e, False), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(pool.float().cpu().numpy(), orc.from_bits(pool_e, False).reshape(pool.shape), rtol=1e-2, atol=1e-2)
    # CUDA-graph style refresh: new block table + lengths, same objects updated in place
    new_ids = block_ids.flip(0).contiguous()
    new_inputs = types.SimpleNamespace(sequence_lengths=seq, kv_cache_kernel_block_id_device=new_ids, is_prefill=False)
    impl.prepare_cuda_graph(new_inputs)
    torch.cuda.synchronize()
    assert np.array_equal(impl.fmha_params.kv_cache_offset.cpu().numpy(), orc.convert_block_table(new_ids.cpu().numpy()))


@pytest.mark.parametrize("fmt", ["gptq", "awq"])
def test_loader_hooks_to_linear_strategy(golden_dir, fmt):
    """AutoGPTQ / AutoAWQ int32 tensors -> B200Impl.preprocess_groupwise_weight_params -> LinearBase strategy."""
    dev = torch.device("cuda")
    g = np.load(os.path.join(golden_dir, f"quant_unpack_{fmt}.npz"))
    impl = device.B200Impl(device="cuda")
    kernel, zs, scales = impl.preprocess_groupwise_weight_params(
        torch.from_numpy(g["qweight"]), torch.from_numpy(g["qzeros"]), torch.from_numpy(g["scales"]), "cuda",
        fmt == "gptq", fmt == "awq", 4)

    class Q:
        def get_method(self):
            return fmt
    assert linear.B200WeightOnlyLinear.can_handle(Q(), kernel.view(torch.int8), scales)
    lin = linear.B200WeightOnlyLinear(kernel, scales, None, None, Q())
    K, N = g["q_packed"].shape[0], g["q_packed"].shape[1] * 2
    x = torch.randn(3, 7, K, device=dev).half()             # leading dims are flattened like F.linear
    y = lin(x)
    assert y.shape == (3, 7, N)
    exp = orc.dequant_gemm(x.reshape(-1, K).cpu().view(torch.int16).numpy().view(np.uint16), "int4", g["q_packed"],
                           scales=g["scales_out"], zeros_x_scales=g["zeros_x_scales"], group=int(g["group"]))
    np.testing.assert_allclose(y.reshape(-1, N).float().cpu().numpy(), orc.from_bits(exp, False), rtol=2e-2, atol=2e-2)


def test_apply_int8_to_linear_strategy(golden_dir):
    dev = torch.device("cuda")
    g = np.load(os.path.join(golden_dir, "quant_int8.npz"))
    w = torch.from_numpy(np.pad(g["weight"], ((0, 64), (0, 0))))     # K 192 -> 256 (kernel needs K % 128 == 0)
    impl = device.B200Impl(device="cuda")
    kernel, scale = impl.apply_int8(w, "cuda")

    class Q:
        def get_method(self):
            return "int8"
    lin = linear.B200WeightOnlyLinear(kernel, scale.half(), None, None, Q())
    x = torch.randn(5, 256, device=dev).half()
    y = lin(x)
    q, s = orc.quantize_int8_per_col(w.numpy())
    exp = orc.dequant_gemm(x.cpu().view(torch.int16).numpy().view(np.uint16), "int8", q, scales=s.astype(np.float16))
    np.testing.assert_allclose(y.float().cpu().numpy(), orc.from_bits(exp, False), rtol=1e-2, atol=1e-2)


def test_f16_linear_matches_f_linear():
    dev = torch.device("cuda")
    w = (torch.randn(256, 384, device=dev) * 0.05).half()            # reference stores [K, N]
    b = torch.randn(384, device=dev).half()
    lin = linear.B200WeightOnlyLinear(w, None, None, b, None)
    x = torch.randn(9, 256, device=dev).half()
    torch.testing.assert_close(lin(x).float(), torch.nn.functional.linear(x.float(), w.t().float(), b.float()),
                               rtol=1e-2, atol=1e-2)
