#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own Python code in this container.

Run here (CPU container, /root/reference mounted):   python oracle/make_golden.py
/root/reference does not exist on the GPU box, so the outputs are committed as small fixtures and this
script is committed beside them.  Nothing in here is product code.

What is imported from the reference (unmodified, loaded from /root/reference by file path):
  * rtp_llm/device/device_impl.py      GpuImpl.preprocess_groupwise_weight_params, apply_int8 /
                                       symmetric_quantize_last_axis_of_batched_matrix, unpack/reverse/pack helpers,
                                       CudaImpl.preprocess_weights_for_mixed_gemm (FT layout, kept as a fixture only)
  * rtp_llm/models_py/modules/factory/attention/cuda_impl/test/atten_test_util.py   attention_prefill_ref
  * .../cuda_impl/test/test_flashinfer_prefill/test_mha_rotary_emb.py   create_cos_sin_cache, apply_rope_reference
                                       (the two functions are extracted by name and executed unmodified)
The modules' unrelated imports (rtp_llm.ops, config, ...) are stubbed because the compiled ops library cannot be
built here (bazel-only).  Expected indexing values follow the reference tests' own
expected-value builders (test_py_flashinfer_mha_decode.py:66-89, trtllm_gen_test.py:305-314), restated here.
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave as a package
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    class _DeviceBase:  # stands in for rtp_llm/device/device_base.py:22 (ctor needs the whole server config)
        def __init__(self):
            pass

    class _W:  # rtp_llm/utils/model_weight.py W: only attribute access at import time
        def __getattr__(self, k):
            return k

    _stub("rtp_llm")
    _stub("rtp_llm.device")
    _stub("rtp_llm.device.device_base", DeviceBase=_DeviceBase, MemInfo=object)
    _stub("rtp_llm.ops")
    _stub("rtp_llm.ops.compute_ops", preprocess_gemm_weight_by_key=None, preprocess_weight_scale=None,
          KVCache=object, PyAttentionInputs=object)
    _stub("rtp_llm.utils")
    _stub("rtp_llm.utils.model_weight", W=_W())
    _stub("rtp_llm.utils.swizzle_utils", swizzle_tensor=None)
    dev = _load("rtp_llm.device.device_impl", f"{REF}/rtp_llm/device/device_impl.py")
    att = _load("ref_atten_test_util",
                f"{REF}/rtp_llm/models_py/modules/factory/attention/cuda_impl/test/atten_test_util.py")
    return dev, att


def gen_quant(dev):
    g = torch.Generator().manual_seed(1)
    K, N, group = 256, 64, 128
    scales = (torch.randn(K // group, N, generator=g).abs() * 0.01 + 1e-3).half()

    class Identity(dev.GpuImpl):
        """Reference GpuImpl with the FT permutation switched off -> the un-permuted tensors."""
        def __init__(self):
            pass

        @property
        def specify_gpu_arch(self):
            return "100"

        def preprocess_weights_for_mixed_gemm(self, tensor, quant_mode, arch=""):
            return tensor

    class FtLayout(dev.CudaImpl):
        def __init__(self):
            pass

        @property
        def specify_gpu_arch(self):
            return "80"

        @property
        def arch(self):
            return 80

    for name, gptq, awq in (("gptq", True, False), ("awq", False, True)):
        if gptq:
            qweight = torch.randint(-2**31, 2**31 - 1, (K // 8, N), generator=g, dtype=torch.int64).int()
        else:
            qweight = torch.randint(-2**31, 2**31 - 1, (K, N // 8), generator=g, dtype=torch.int64).int()
        qzeros = torch.randint(-2**31, 2**31 - 1, (K // group, N // 8), generator=g, dtype=torch.int64).int()
        qp, zs, sc = Identity().preprocess_groupwise_weight_params(qweight.clone(), qzeros.clone(), scales.clone(),
                                                                   "cpu", gptq, awq, 4)
        ft, _, _ = FtLayout().preprocess_groupwise_weight_params(qweight.clone(), qzeros.clone(), scales.clone(),
                                                                 "cpu", gptq, awq, 4)
        np.savez_compressed(os.path.join(OUT, f"quant_unpack_{name}.npz"),
                            qweight=qweight.numpy(), qzeros=qzeros.numpy(), scales=scales.numpy(),
                            q_packed=qp.numpy().view(np.uint8), zeros_x_scales=zs.numpy(), scales_out=sc.numpy(),
                            ft_layout=ft.numpy().view(np.uint8), group=group)
        print(name, "q_packed", tuple(qp.shape), qp.dtype, "zs", tuple(zs.shape))

    # INT8 per-column (apply_int8 -> symmetric_quantize_last_axis_of_batched_matrix), device_impl.py:183-222
    w = (torch.randn(192, 96, generator=g) * 0.02).float()
    q8, s8 = Identity().apply_int8(w.clone(), "cpu")
    np.savez_compressed(os.path.join(OUT, "quant_int8.npz"), weight=w.numpy(), q=q8.numpy(), scale=s8.numpy())
    print("int8", tuple(q8.shape), q8.dtype, tuple(s8.shape), s8.dtype)


def gen_quant_w8(dev):
    """8-bit group-wise checkpoints: the is_int8 branch of preprocess_groupwise_weight_params (device_impl.py:256-258,
    unpack_int32_into_int16 :147-149): one byte per weight, zero shift 128, no nibble packing."""
    g = torch.Generator().manual_seed(8)
    K, N, group = 256, 64, 128
    scales = (torch.randn(K // group, N, generator=g).abs() * 0.01 + 1e-3).half()

    class Identity(dev.GpuImpl):
        def __init__(self):
            pass

        @property
        def specify_gpu_arch(self):
            return "100"

        def preprocess_weights_for_mixed_gemm(self, tensor, quant_mode, arch=""):
            return tensor

    for name, gptq, awq in (("gptq", True, False), ("awq", False, True)):
        if gptq:
            qweight = torch.randint(-2**31, 2**31 - 1, (K // 4, N), generator=g, dtype=torch.int64).int()
        else:
            qweight = torch.randint(-2**31, 2**31 - 1, (K, N // 4), generator=g, dtype=torch.int64).int()
        qzeros = torch.randint(-2**31, 2**31 - 1, (K // group, N // 4), generator=g, dtype=torch.int64).int()
        q8, zs, sc = Identity().preprocess_groupwise_weight_params(qweight.clone(), qzeros.clone(), scales.clone(),
                                                                   "cpu", gptq, awq, 8)
        assert q8.dtype == torch.int8 and tuple(q8.shape) == (K, N)
        np.savez_compressed(os.path.join(OUT, f"quant_unpack8_{name}.npz"),
                            qweight=qweight.numpy(), qzeros=qzeros.numpy(), scales=scales.numpy(),
                            q=q8.numpy(), zeros_x_scales=zs.numpy(), scales_out=sc.numpy(), group=group)
        print(name, "w8 q", tuple(q8.shape), q8.dtype, "zs", tuple(zs.shape), zs.dtype)


def gen_attention(att):
    """Decode attention goldens through the reference's torch oracle attention_prefill_ref
    (atten_test_util.py:55-116): the decode query is placed at the last position of each sequence (other
    query rows are zero); with causal=True the last row attends to the whole sequence = decode semantics."""
    cases = [
        # name, seed, dist, Hq, Hkv, D, page, lens (KV length INCLUDING the new token)
        ("attn_p16_gqa4", 42, "randn", 8, 2, 128, 16, [10, 20, 65, 130]),          # test_xqa.py:368-449 style
        ("attn_p64_gqa8", 25536, "uniform", 16, 2, 128, 64, [2, 129, 255, 63]),    # trtllm_gen_test.py:39-244 style
        ("attn_p32_mha", 42, "randn", 4, 4, 128, 32, [64, 65, 1]),
    ]
    for name, seed, dist, Hq, Hkv, D, T, lens in cases:
        g = torch.Generator().manual_seed(seed)
        B = len(lens)
        M = max(math.ceil(L / T) for L in lens)
        npages = sum(math.ceil(L / T) for L in lens) + 1  # block 0 reserved as the null block
        def rnd(*shape):
            if dist == "randn":
                return torch.randn(*shape, generator=g)
            return torch.rand(*shape, generator=g) * 2 - 1
        pool = rnd(npages, 2, Hkv, T, D).half()
        q = rnd(B, Hq, D).half()
        perm = (torch.randperm(npages - 1, generator=g) + 1).tolist()
        block_ids = torch.zeros(B, M, dtype=torch.int32)
        it = iter(perm)
        for b, L in enumerate(lens):
            for j in range(math.ceil(L / T)):
                block_ids[b, j] = next(it)
        # un-page into the [tokens, heads, dim] tensors the reference oracle takes
        ks, vs, qs = [], [], []
        for b, L in enumerate(lens):
            pages = block_ids[b, : math.ceil(L / T)].long()
            k = pool[pages, 0].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :L].permute(1, 0, 2)
            v = pool[pages, 1].permute(1, 0, 2, 3).reshape(Hkv, -1, D)[:, :L].permute(1, 0, 2)
            qq = torch.zeros(L, Hq, D, dtype=torch.half)
            qq[-1] = q[b]
            ks.append(k); vs.append(v); qs.append(qq)
        out = att.attention_prefill_ref(torch.cat(qs).float(), torch.cat(ks).float(), torch.cat(vs).float(),
                                        torch.tensor(lens), Hq, Hkv, D, causal=True)
        ends = np.cumsum(lens) - 1
        expect = out[torch.tensor(ends)].reshape(B, Hq * D).float().numpy()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), q=q.numpy(), kv_pool=pool.numpy(),
                            block_ids=block_ids.numpy(), sequence_lengths=np.array([L - 1 for L in lens], np.int32),
                            expect=expect, head_num=Hq, kv_head_num=Hkv, head_dim=D, tokens_per_block=T)
        print(name, "expect", expect.shape, float(np.abs(expect).max()))


def _ref_functions(path, names):
    """Extract top-level functions of a reference file BY NAME and execute them unmodified (the file's own imports need
    CUDA / the compiled ops library; the functions themselves are pure torch)."""
    import ast
    from typing import Tuple
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "Tuple": Tuple, "math": math}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def gen_rope():
    """RoPE (RopeStyle::Base, NeoX / non-interleaved pairing) goldens from the reference's pure-torch implementation:
    create_cos_sin_cache + apply_rope_reference, test_flashinfer_prefill/test_mha_rotary_emb.py:47-81,121-165 -- the oracle
    its own fused-rope kernel tests compare against (rtol = atol = 1e-2, :506-507). Decode: one token per sequence at
    position sequence_lengths[b]; computed in fp32 and rounded to fp16 exactly as the reference test does (:272-276)."""
    path = f"{REF}/rtp_llm/models_py/modules/factory/attention/cuda_impl/test/test_flashinfer_prefill/test_mha_rotary_emb.py"
    create_cos_sin_cache, apply_rope_reference = _ref_functions(path, ["create_cos_sin_cache", "apply_rope_reference"])
    g = torch.Generator().manual_seed(7)
    for name, B, Hq, Hkv, D, base, max_pos in (("rope_base10000", 5, 8, 2, 128, 10000.0, 4096),
                                               ("rope_base500000", 4, 4, 4, 128, 500000.0, 8192),
                                               ("rope_d64", 3, 4, 1, 64, 10000.0, 2048)):
        qkv = torch.randn(B, (Hq + 2 * Hkv) * D, generator=g).half()
        positions = torch.randint(0, max_pos, (B,), generator=g)
        positions[0] = 0
        positions[-1] = max_pos - 1
        cache = create_cos_sin_cache(D, max_pos, base, device="cpu")
        q = qkv[:, : Hq * D].reshape(B, Hq, D)
        k = qkv[:, Hq * D:(Hq + Hkv) * D].reshape(B, Hkv, D)
        q_ref, k_ref = apply_rope_reference(q.float(), k.float(), cache.float(), positions)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), qkv=qkv.numpy(), positions=positions.numpy().astype(np.int32),
                            q_rope=q_ref.half().numpy(), k_rope=k_ref.half().numpy(), head_num=Hq, kv_head_num=Hkv,
                            head_dim=D, rope_base=np.float32(base))
        print(name, tuple(q_ref.shape), tuple(k_ref.shape))


def gen_rope_cache_styles():
    """Cache path of the decode rope op (Base with linear scale, Yarn): the cos/sin table is built exactly as
    cpp/model_utils/RopeCache.cc:16-85 builds it (the same torch calls, on CPU), re-laid out as [pos, (cos half | sin half)]
    for the reference's pure-torch apply_rope_reference (test_mha_rotary_emb.py:121-165), which produces the expected values."""
    path = f"{REF}/rtp_llm/models_py/modules/factory/attention/cuda_impl/test/test_flashinfer_prefill/test_mha_rotary_emb.py"
    (apply_rope_reference,) = _ref_functions(path, ["apply_rope_reference"])

    def base_cache(dim, theta, scale, max_pos):          # RopeCache.cc:16-44
        inv_freq = 1.0 / torch.pow(torch.tensor(float(theta)), torch.arange(0, dim, 2).float() / dim)
        t = torch.arange(int(max_pos * scale)).float()
        t.div_(scale)
        freqs = torch.outer(t, inv_freq)
        return freqs.cos(), freqs.sin()

    def yarn_cache(dim, theta, scale, max_pos, beta_slow, beta_fast, extrapolation_factor, mscale):   # RopeCache.cc:46-85
        pos_freqs = torch.pow(torch.tensor(float(theta)), torch.arange(0, dim, 2).float() / dim)
        inv_e, inv_i = 1.0 / pos_freqs, 1.0 / (scale * pos_freqs)
        corr = lambda nrot: float(dim * math.log(max_pos / (nrot * 2.0 * math.pi))) / (2.0 * math.log(theta))
        low = float(max(0, int(math.floor(corr(beta_slow)))))
        high = float(min(dim - 1, int(math.ceil(corr(beta_fast)))))
        if abs(low - high) < 1e-6:
            high += 0.001
        ramp = torch.clamp((torch.arange(dim // 2).float() - low) / (high - low), 0, 1)
        mask = (1.0 - ramp) * extrapolation_factor
        inv_freq = inv_i * (1.0 - mask) + inv_e * mask
        freqs = torch.outer(torch.arange(int(max_pos * scale)).float(), inv_freq)
        return freqs.cos() * mscale, freqs.sin() * mscale

    g = torch.Generator().manual_seed(11)
    cases = [("rope_cache_base_scale2", dict(style=1, dim=128, base=10000.0, scale=2.0, factor1=1.0, factor2=1.0, max_pos=2048,
                                             extrapolation_factor=1.0, mscale=1.0), base_cache(128, 10000, 2.0, 2048)),
             ("rope_cache_yarn", dict(style=5, dim=128, base=10000.0, scale=4.0, factor1=1.0, factor2=32.0, max_pos=1024,
                                      extrapolation_factor=1.0, mscale=1.13), yarn_cache(128, 10000, 4.0, 1024, 1, 32, 1.0, 1.13))]
    for name, cfg, (cos, sin) in cases:
        B, Hq, Hkv, D = 4, 4, 2, 128
        qkv = torch.randn(B, (Hq + 2 * Hkv) * D, generator=g).half()
        npos = cos.shape[0]
        positions = torch.randint(1, npos, (B,), generator=g)
        positions[-1] = npos - 1
        cache = torch.cat([cos, sin], dim=-1)
        q = qkv[:, : Hq * D].reshape(B, Hq, D)
        k = qkv[:, Hq * D:(Hq + Hkv) * D].reshape(B, Hkv, D)
        q_ref, k_ref = apply_rope_reference(q.float(), k.float(), cache.float(), positions)
        inter = torch.stack([cos, sin], dim=-1).reshape(npos, -1).contiguous()      # getRopeCache(interleave = true) layout
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), qkv=qkv.numpy(), positions=positions.numpy().astype(np.int32),
                            q_rope=q_ref.half().numpy(), k_rope=k_ref.half().numpy(), cache_rows=inter[positions].numpy(), cache_positions=npos,
                            head_num=Hq, kv_head_num=Hkv, head_dim=D, **{f"cfg_{k}": v for k, v in cfg.items()})
        print(name, tuple(inter.shape))


def gen_indexing():
    """Expected values built the way the reference tests build them."""
    rng = np.random.default_rng(42)
    # (1) block table -> [B,2,M] offsets, trtllm_gen_test.py:305-314 (_reference_kv_offset)
    block_id = rng.integers(0, 512, size=(4, 8), dtype=np.int32)
    kv_offset = np.zeros((4, 2, 8), np.int32)
    for b in range(4):
        for m in range(8):
            kv_offset[b, 0, m] = block_id[b, m] * 2
            kv_offset[b, 1, m] = block_id[b, m] * 2 + 1
    # (2) flashinfer plan, test_py_flashinfer_mha_decode.py:66-89 (sequential block ids, lens are KV lengths incl. new token)
    plans = {}
    for tag, lens, T in (("a", [64, 128, 256, 512], 64), ("b", [10, 20], 16), ("c", [65], 64), ("d", [1, 16, 17, 33], 16)):
        indptr, indices, last = [0], [], []
        off = 0
        for L in lens:
            nb = math.ceil(L / T)
            indptr.append(indptr[-1] + nb)
            indices += [off + j for j in range(nb)]
            last.append(L % T or T)
            off += nb
        plans[f"plan_{tag}_lens"] = np.array(lens, np.int32)
        plans[f"plan_{tag}_T"] = np.int32(T)
        plans[f"plan_{tag}_indptr"] = np.array(indptr, np.int32)
        plans[f"plan_{tag}_indices"] = np.array(indices, np.int32)
        plans[f"plan_{tag}_last"] = np.array(last, np.int32)
    np.savez_compressed(os.path.join(OUT, "indexing.npz"), block_id=block_id, kv_offset=kv_offset, **plans)
    print("indexing ok")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    dev, att = load_reference()
    only = set(sys.argv[1:])          # e.g. `make_golden.py quant_w8` regenerates one family; no argument = all

    def want(name):
        return not only or name in only
    if want("quant"):
        gen_quant(dev)
    if want("quant_w8"):
        gen_quant_w8(dev)
    if want("attention"):
        gen_attention(att)
    if want("rope"):
        gen_rope()
    if want("rope_cache"):
        gen_rope_cache_styles()
    if want("indexing"):
        gen_indexing()
