"""Synthetic AutoGPTQ / AutoAWQ checkpoints for the loader tests: random 4-bit weights + zeros + scales packed the way the two
tools store them, written with safetensors, plus the dense weight they mean.

AutoGPTQ:  qweight int32 [K/8, N]  (8 rows per word, row k in bits 4*(k%8)),  qzeros int32 [K/g, N/8] storing z - 1 (column n in
           bits 4*(n%8)), scales fp16 [K/g, N];  W[k][n] = (q - (qz + 1)) * s
AutoAWQ:   qweight int32 [K, N/8] with the 8 columns of a word in the order 0,2,4,6,1,3,5,7,  qzeros int32 [K/g, N/8] same order,
           scales fp16 [K/g, N];  W[k][n] = (q - z) * s
(the conventions device_impl.py:242-300 undoes: reverse_awq_order for AWQ, the "- GPTQ_FLAG" for GPTQ's stored z - 1)."""
import numpy as np
import torch

AWQ_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]


def _pack_cols(vals, order, bits=4):
    """vals [R, C] in 0..2^bits-1 -> int32 [R, C*bits/32]: field j of a run of 8 columns = column order[j] of the run (4-bit: one
    word per run; 8-bit: two words, the reference applies reverse_awq_order over runs of 8 either way, device_impl.py:163-171)."""
    R, C = vals.shape
    v = vals.reshape(R, C // 8, 8).astype(np.uint64)
    out = np.zeros((R, C // 8), np.uint64)
    for j, src in enumerate(order):
        out |= v[:, :, src] << np.uint64(bits * j)
    if bits == 4:
        return out.astype(np.uint32).view(np.int32)
    return np.ascontiguousarray(out).view(np.uint32).reshape(R, C // 4).view(np.int32)     # little endian: low word first


def _pack_rows(vals, bits=4):
    """vals [R, C] -> int32 [R*bits/32, C], field j of a word = row per*i + j."""
    per = 32 // bits
    R, C = vals.shape
    v = vals.reshape(R // per, per, C).astype(np.uint32)
    out = np.zeros((R // per, C), np.uint32)
    for j in range(per):
        out |= v[:, j, :] << (bits * j)
    return out.view(np.int32)


def make_layer(rng, method, K, N, group=128, bits=4):
    top = 1 << bits
    q = rng.integers(0, top, (K, N))
    z = rng.integers(1 if method == "gptq" else 0, top, (K // group, N))
    s = (np.abs(rng.standard_normal((K // group, N))) * (0.01 if bits == 4 else 6e-4) + (1e-3 if bits == 4 else 6e-5)).astype(np.float16)
    dense = ((q - np.repeat(z, group, 0)).astype(np.float32) * np.repeat(s.astype(np.float32), group, 0))
    if method == "gptq":
        t = dict(qweight=_pack_rows(q, bits), qzeros=_pack_cols(z - 1, list(range(8)), bits), scales=s)
    else:
        t = dict(qweight=_pack_cols(q, AWQ_ORDER, bits), qzeros=_pack_cols(z, AWQ_ORDER, bits), scales=s)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in t.items()}, dense


def write_checkpoint(path, rng, method, hidden, heads, kv_heads, head_dim, inter, group=128, bits=4):
    """One decoder layer in HF naming; returns the dense [K, N] weights by logical name."""
    from safetensors.torch import save_file
    shapes = dict(q=(hidden, heads * head_dim), k=(hidden, kv_heads * head_dim), v=(hidden, kv_heads * head_dim),
                  o=(heads * head_dim, hidden), gate=(hidden, inter), up=(hidden, inter), down=(inter, hidden))
    hf = dict(q="self_attn.q_proj", k="self_attn.k_proj", v="self_attn.v_proj", o="self_attn.o_proj", gate="mlp.gate_proj",
              up="mlp.up_proj", down="mlp.down_proj")
    tensors, dense, names = {}, {}, {}
    for key, (K, N) in shapes.items():
        names[key] = f"model.layers.0.{hf[key]}"
        if method in ("gptq", "awq"):
            t, d = make_layer(rng, method, K, N, group, bits)
            for suffix, val in t.items():
                tensors[f"{names[key]}.{suffix}"] = val
        else:
            d = (rng.standard_normal((K, N)) * 0.02).astype(np.float32)
            tensors[f"{names[key]}.weight"] = torch.from_numpy(np.ascontiguousarray(d.T)).half()
            d = tensors[f"{names[key]}.weight"].float().numpy().T
        dense[key] = d
    save_file(tensors, path)
    return names, dense
