"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see decode_oracle.c header).

Two layers:
  * ctypes bindings to oracle/liboracle.so (the C restatement, also the timed CPU arm), and
  * small pure-numpy restatements of the same functions used to cross-check the C code on tiny cases.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[ctypes.CDLL] = None


def build() -> str:
    """Compile oracle/liboracle.so (gcc, OpenMP). Building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return os.path.join(_HERE, "liboracle.so")


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_num_threads.restype = ctypes.c_int
        _LIB.oracle_paged_attn_plan.restype = ctypes.c_int
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"], "oracle expects contiguous arrays"
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ----------------------------------------------------------------------------------------------
# 16-bit helpers: tensors cross the oracle boundary as uint16 bit patterns
# ----------------------------------------------------------------------------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    lsb = (u >> 16) & 1
    return ((u + 0x7FFF + lsb) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def to_bits(x: np.ndarray, is_bf16: bool) -> np.ndarray:
    if x.dtype == np.uint16:
        return np.ascontiguousarray(x)
    if is_bf16:
        return f32_to_bf16_bits(x)
    return np.ascontiguousarray(x.astype(np.float16)).view(np.uint16)


def from_bits(b: np.ndarray, is_bf16: bool) -> np.ndarray:
    return bf16_bits_to_f32(b) if is_bf16 else b.view(np.float16).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# Indexing
# ----------------------------------------------------------------------------------------------
def convert_block_table(block_ids: np.ndarray) -> np.ndarray:
    """kv_cache_kernels.cu:49-63. [B,M] int32 -> [B,1,2,M] int32 (K page 2*id, V page 2*id+1)."""
    block_ids = np.ascontiguousarray(block_ids, dtype=np.int32)
    B, M = block_ids.shape
    out = np.empty((B, 1, 2, M), dtype=np.int32)
    lib().oracle_convert_block_table(_p(out), _p(block_ids), B, M)
    return out


def convert_block_table_np(block_ids: np.ndarray) -> np.ndarray:
    b = np.asarray(block_ids, dtype=np.int32)
    return np.stack([b * 2, b * 2 + 1], axis=1)[:, None, :, :].astype(np.int32)


def paged_attn_plan(sequence_lengths, block_ids, tokens_per_block, input_lengths=None, prefix_lengths=None):
    """mha_paged_attn_plan.cu:28-97. Returns dict of last_page_len, page_indptr, page_indice, batch_indice, positions."""
    B = len(sequence_lengths) if prefix_lengths is None else len(prefix_lengths)
    block_ids = np.ascontiguousarray(block_ids, dtype=np.int32)
    M = block_ids.shape[1]
    seq = np.ascontiguousarray(sequence_lengths, dtype=np.int32) if sequence_lengths is not None else None
    inl = np.ascontiguousarray(input_lengths, dtype=np.int32) if input_lengths is not None else None
    pre = np.ascontiguousarray(prefix_lengths, dtype=np.int32) if prefix_lengths is not None else None
    ntok = int(inl.sum()) if pre is not None else B
    last = np.zeros(B, np.int32)
    indptr = np.zeros(B + 1, np.int32)
    indice = np.zeros(B * M, np.int32)
    bidx = np.zeros(max(ntok, 1), np.int32)
    pos = np.zeros(max(ntok, 1), np.int32)
    total = lib().oracle_paged_attn_plan(
        _p(inl) if inl is not None else None,
        _p(seq) if seq is not None else None,
        _p(pre) if pre is not None else None,
        _p(block_ids), B, M, int(tokens_per_block), _p(last), _p(indptr), _p(indice), _p(bidx), _p(pos))
    return dict(last_page_len=last, page_indptr=indptr, page_indice=indice[:total], batch_indice=bidx[:ntok],
                positions=pos[:ntok])


# ----------------------------------------------------------------------------------------------
# Paged decode attention
# ----------------------------------------------------------------------------------------------
def paged_decode_attn(q_bits, kv_pool_bits, page_list, sequence_lengths, head_num, kv_head_num, head_dim,
                      tokens_per_block, is_bf16=False, q_scale=1.0) -> np.ndarray:
    """q [B,Hq,D] bits, kv_pool [P,2,Hkv,T,D] bits, page_list [B,1,2,M] int32, sequence_lengths [B] (past len).
    Returns out bits [B, Hq*D]."""
    q_bits = np.ascontiguousarray(q_bits, dtype=np.uint16)
    kv_pool_bits = np.ascontiguousarray(kv_pool_bits, dtype=np.uint16)
    page_list = np.ascontiguousarray(page_list, dtype=np.int32)
    seq = np.ascontiguousarray(sequence_lengths, dtype=np.int32)
    B = q_bits.shape[0]
    M = page_list.shape[-1]
    out = np.empty((B, head_num * head_dim), dtype=np.uint16)
    lib().oracle_paged_decode_attn(_p(q_bits), int(is_bf16), _p(out), head_num, kv_head_num, head_dim, B, M,
                                   tokens_per_block, _p(kv_pool_bits), _p(page_list), _p(seq),
                                   ctypes.c_float(q_scale))
    return out


def paged_decode_attn_np(q, kv_pool, block_ids, sequence_lengths, tokens_per_block) -> np.ndarray:
    """Tiny pure-numpy restatement (float64, no P rounding) for cross-checks. q [B,Hq,D] float, kv_pool [P,2,Hkv,T,D]."""
    B, Hq, D = q.shape
    Hkv = kv_pool.shape[2]
    g = Hq // Hkv
    out = np.zeros((B, Hq, D), np.float64)
    for b in range(B):
        L = int(sequence_lengths[b]) + 1
        npages = (L + tokens_per_block - 1) // tokens_per_block
        pages = [int(block_ids[b][j]) for j in range(npages)]
        k = np.concatenate([kv_pool[p, 0] for p in pages], axis=1)[:, :L]  # [Hkv, L, D]
        v = np.concatenate([kv_pool[p, 1] for p in pages], axis=1)[:, :L]
        for h in range(Hq):
            s = (k[h // g].astype(np.float64) @ q[b, h].astype(np.float64)) / np.sqrt(D)
            p = np.exp(s - s.max())
            out[b, h] = (p[:, None] * v[h // g].astype(np.float64)).sum(0) / p.sum()
    return out.reshape(B, Hq * D)


# ----------------------------------------------------------------------------------------------
# Weight-only quantisation
# ----------------------------------------------------------------------------------------------
def unpack_groupwise_int4(qweight, qzeros, scales_f16, group: int, is_gptq: bool):
    """device_impl.py:242-300 (4-bit). Returns (q_packed uint8 [K,N/2], zeros_x_scales f16 [K/g,N], scales f16)."""
    qweight = np.ascontiguousarray(qweight, dtype=np.int32)
    qzeros = np.ascontiguousarray(qzeros, dtype=np.int32)
    scales = np.ascontiguousarray(scales_f16, dtype=np.float16)
    if is_gptq:
        K, N = qweight.shape[0] * 8, qweight.shape[1]
    else:
        K, N = qweight.shape[0], qweight.shape[1] * 8
    qp = np.zeros((K, N // 2), np.uint8)
    zs = np.zeros((K // group, N), np.float16)
    lib().oracle_unpack_groupwise_int4(_p(qweight), _p(qzeros), _p(scales.view(np.uint16)), K, N, group,
                                       int(is_gptq), _p(qp), _p(zs.view(np.uint16)))
    return qp, zs, scales


def unpack_groupwise_int8(qweight, qzeros, scales_f16, group: int, is_gptq: bool):
    """device_impl.py:242-300 with weight_bits == 8 (is_int8: zero shift 128, one byte per weight, :147-149,256-258).
    GPTQ qweight int32 [K/4, N] (4 bytes along K, low byte first), AWQ qweight int32 [K, N/4] (bytes along N, then the
    reference applies reverse_awq_order over runs of 8 columns, :163-171,267), qzeros int32 [K/g, N/4].
    Returns (q_s int8 [K,N], zeros_x_scales f16 [K/g,N], scales f16)."""
    qweight = np.ascontiguousarray(qweight, dtype=np.int32)
    qzeros = np.ascontiguousarray(qzeros, dtype=np.int32)
    scales = np.ascontiguousarray(scales_f16, dtype=np.float16)

    def awq_order(t):                       # out[..., 8g + 2j + i] = in[..., 8g + 4i + j]
        return t.reshape(-1, 2, 4).transpose(0, 2, 1).reshape(t.shape)
    if is_gptq:
        u = np.ascontiguousarray(qweight.T).view(np.uint8).T.astype(np.int16)      # [K, N]
    else:
        u = awq_order(qweight.view(np.uint8).astype(np.int16))
    q = (u - 128).astype(np.int8)
    z = qzeros.view(np.uint8).astype(np.int16)
    if not is_gptq:
        z = awq_order(z)
    zf = (-z + 128 - (1 if is_gptq else 0)).astype(np.float16)                      # |.| <= 128: exact in fp16
    zs = (zf.astype(np.float32) * scales.astype(np.float32)).astype(np.float16)     # one rounding, as torch's half multiply
    return np.ascontiguousarray(q), zs, scales


def quantize_int8_per_col(w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """device_impl.py:183-202 int8 branch. w [K,N] float -> (q int8 [K,N], scale fp32 [N])."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    K, N = w.shape
    q = np.empty((K, N), np.int8)
    s = np.empty(N, np.float32)
    lib().oracle_quantize_int8_per_col(_p(w), K, N, _p(q), _p(s))
    return q, s


def unpack_int4_signed(q_packed: np.ndarray) -> np.ndarray:
    """[K,N/2] uint8 -> q_s int8 [K,N] in [-8,7] (low nibble = even column, device_impl.py:204-209)."""
    lo = (q_packed & 0xF).astype(np.int8)
    hi = (q_packed >> 4).astype(np.int8)
    lo = np.where(lo >= 8, lo - 16, lo)
    hi = np.where(hi >= 8, hi - 16, hi)
    out = np.empty((q_packed.shape[0], q_packed.shape[1] * 2), np.int8)
    out[:, 0::2] = lo
    out[:, 1::2] = hi
    return out


def dequant_np(fmt: str, weight, scales=None, zeros_x_scales=None, group: int = 128, is_bf16=False) -> np.ndarray:
    """W' [K,N] float32 per SURVEY section 8 a9/a10 (rounded to the activation type)."""
    def rnd(x):
        return from_bits(to_bits(x, is_bf16), is_bf16)
    if fmt == "f16":
        return np.asarray(weight, np.float32)
    if fmt == "int8":
        return rnd(weight.astype(np.float32) * np.asarray(scales, np.float32)[None, :])
    if fmt == "int4":
        qs = unpack_int4_signed(weight).astype(np.float32)
        s = np.repeat(np.asarray(scales, np.float32), group, axis=0)
        z = np.repeat(np.asarray(zeros_x_scales, np.float32), group, axis=0)
        return rnd(qs * s + z)
    if fmt == "int8g":
        s = np.repeat(np.asarray(scales, np.float32), group, axis=0)
        z = np.repeat(np.asarray(zeros_x_scales, np.float32), group, axis=0)
        return rnd((np.asarray(weight).astype(np.float64) * s + z).astype(np.float32))   # fma: one rounding to fp32, then to the element type
    raise ValueError(fmt)


_FMT = {"f16": 0, "int8": 1, "int4": 2, "int8g": 3}


def dequant_gemm(x_bits, fmt: str, weight, scales=None, zeros_x_scales=None, group: int = 128, bias=None,
                 is_bf16=False, fast=False) -> np.ndarray:
    """Y = X . W' (+bias). x_bits [B,K] uint16; weight per fmt (f16: uint16 bits [K,N]; int8: int8 [K,N];
    int4: uint8 [K,N/2]; int8g: int8 [K,N]); scales / zeros as uint16 bits (int8: [N]; int4 / int8g: [K/g,N]).
    Returns bits [B,N]."""
    x_bits = np.ascontiguousarray(x_bits, dtype=np.uint16)
    B, K = x_bits.shape
    weight = np.ascontiguousarray(weight)
    N = weight.shape[1] * (2 if fmt == "int4" else 1)
    y = np.empty((B, N), np.uint16)
    sc = to_bits(np.asarray(scales), is_bf16) if scales is not None else None
    zs = to_bits(np.asarray(zeros_x_scales), is_bf16) if zeros_x_scales is not None else None
    bs = to_bits(np.asarray(bias), is_bf16) if bias is not None else None
    if fast:
        assert bias is None
        lib().oracle_dequant_gemm_fast(_p(x_bits), int(is_bf16), B, K, N, _FMT[fmt], _p(weight),
                                       _p(sc) if sc is not None else None, _p(zs) if zs is not None else None,
                                       int(group), _p(y))
    else:
        lib().oracle_dequant_gemm(_p(x_bits), int(is_bf16), B, K, N, _FMT[fmt], _p(weight),
                                  _p(sc) if sc is not None else None, _p(zs) if zs is not None else None,
                                  int(group), _p(bs) if bs is not None else None, _p(y))
    return y


# ----------------------------------------------------------------------------------------------
# Glue ops
# ----------------------------------------------------------------------------------------------
def add_rmsnorm(x_bits, residual_bits, gamma_bits, eps, is_bf16=False):
    x_bits = np.ascontiguousarray(x_bits, np.uint16)
    rows, hidden = x_bits.shape
    y = np.empty_like(x_bits)
    res = np.ascontiguousarray(residual_bits, np.uint16).copy() if residual_bits is not None else None
    lib().oracle_add_rmsnorm(_p(x_bits), _p(res) if res is not None else None,
                             _p(np.ascontiguousarray(gamma_bits, np.uint16)), _p(y), int(is_bf16), rows, hidden,
                             ctypes.c_float(eps), int(res is not None))
    return y, res


def silu_and_mul(gate_up_bits, is_bf16=False):
    g = np.ascontiguousarray(gate_up_bits, np.uint16)
    rows, two_inter = g.shape
    y = np.empty((rows, two_inter // 2), np.uint16)
    lib().oracle_silu_and_mul(_p(g), _p(y), int(is_bf16), rows, two_inter // 2)
    return y


def rope_append(qkv_bits, kv_pool_bits, page_list, sequence_lengths, head_num, kv_head_num, head_dim,
                tokens_per_block, rope_base, is_bf16=False):
    """Returns (q_out bits [B,Hq*D], updated kv_pool bits copy)."""
    qkv = np.ascontiguousarray(qkv_bits, np.uint16)
    pool = np.ascontiguousarray(kv_pool_bits, np.uint16).copy()
    page_list = np.ascontiguousarray(page_list, np.int32)
    seq = np.ascontiguousarray(sequence_lengths, np.int32)
    B = qkv.shape[0]
    q_out = np.empty((B, head_num * head_dim), np.uint16)
    lib().oracle_rope_append(_p(qkv), _p(q_out), _p(pool), _p(page_list), _p(seq), int(is_bf16), B, head_num,
                             kv_head_num, head_dim, page_list.shape[-1], tokens_per_block, ctypes.c_float(rope_base))
    return q_out, pool


def argmax(logits: np.ndarray) -> np.ndarray:
    logits = np.ascontiguousarray(logits, np.float32)
    out = np.empty(logits.shape[0], np.int32)
    lib().oracle_argmax(_p(logits), logits.shape[0], logits.shape[1], _p(out))
    return out


def sample(logits, top_k, top_p, uniform, temperature=None, history=None, hist_len=None, repetition=None, presence=None,
           frequency=None, process=None):
    """sampleGreedy's path on fp32 logits [B, V]. Returns (tokens, token_prob, probs after softmax, renormalised kept probs)."""
    lg = np.ascontiguousarray(logits, np.float32).copy()
    B, V = lg.shape
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    hist = None if history is None else np.ascontiguousarray(history, np.int32)
    hl = None if hist_len is None else np.ascontiguousarray(hist_len, np.int32)
    tk, tp, u = np.ascontiguousarray(top_k, np.int32), f32(top_p), f32(uniform)
    t, rp, pp, fp = f32(temperature), f32(repetition), f32(presence), f32(frequency)
    pr = None if process is None else np.ascontiguousarray(process, np.uint8)
    tok = np.empty(B, np.int32)
    tprob = np.empty(B, np.float32)
    probs = np.empty((B, V), np.float32)
    P = lambda a: None if a is None else _p(a)
    lib().oracle_sample(_p(lg), B, V, P(hist), P(hl), hist.shape[1] if hist is not None else 0, P(t), P(rp), P(pp), P(fp), _p(tk),
                        _p(tp), _p(u), P(pr), _p(tok), _p(tprob), _p(probs))
    return tok, tprob, lg, probs


class RopeConfig(ctypes.Structure):
    _fields_ = [("style", ctypes.c_int), ("dim", ctypes.c_int), ("base", ctypes.c_float), ("scale", ctypes.c_float),
                ("factor1", ctypes.c_float), ("factor2", ctypes.c_float), ("max_pos", ctypes.c_int),
                ("extrapolation_factor", ctypes.c_float), ("mscale", ctypes.c_float)]


def rope_append_ex(qkv_bits, kv_pool_bits, page_list, sequence_lengths, head_num, kv_head_num, head_dim, tokens_per_block, cfg,
                   bias_bits=None, position_ids=None, cos_sin_cache=None, use_logn=False, is_bf16=False):
    """The full decode rope contract (b200_rope_append_ex). cfg: dict of RopeConfig fields. Returns (q_out bits, pool copy)."""
    qkv = np.ascontiguousarray(qkv_bits, np.uint16)
    pool = np.ascontiguousarray(kv_pool_bits, np.uint16).copy()
    page_list = np.ascontiguousarray(page_list, np.int32)
    seq = np.ascontiguousarray(sequence_lengths, np.int32)
    B = qkv.shape[0]
    q_out = np.empty((B, head_num * head_dim), np.uint16)
    rc = RopeConfig(**cfg)
    bias = None if bias_bits is None else np.ascontiguousarray(bias_bits, np.uint16)
    pid = None if position_ids is None else np.ascontiguousarray(position_ids, np.int32)
    cache = None if cos_sin_cache is None else np.ascontiguousarray(cos_sin_cache, np.float32)
    P = lambda a: None if a is None else _p(a)
    lib().oracle_rope_append_ex(_p(qkv), P(bias), _p(q_out), _p(pool), _p(page_list), _p(seq), P(pid), P(cache),
                                cache.shape[0] if cache is not None else 0, ctypes.byref(rc), int(use_logn), int(is_bf16), B, head_num,
                                kv_head_num, head_dim, page_list.shape[-1], tokens_per_block)
    return q_out, pool


def rope_cache_base(dim, theta, scale, max_pos):
    """genBaseCache (cpp/model_utils/RopeCache.cc:16-44), interleave = true: fp32 [positions, dim] = (cos, sin) pairs."""
    inv_freq = (1.0 / np.power(np.float32(theta), np.arange(0, dim, 2, dtype=np.float32) / np.float32(dim))).astype(np.float32)
    t = (np.arange(int(max_pos * scale), dtype=np.float32) / np.float32(scale)).astype(np.float32)
    freqs = np.outer(t, inv_freq).astype(np.float32)
    return np.stack([np.cos(freqs), np.sin(freqs)], axis=-1).reshape(freqs.shape[0], -1).astype(np.float32)


def rope_cache_yarn(dim, theta, scale, max_pos, beta_slow, beta_fast, extrapolation_factor, mscale):
    """genYarnCache (RopeCache.cc:46-85), interleave = true."""
    pos_freqs = np.power(np.float32(theta), np.arange(0, dim, 2, dtype=np.float32) / np.float32(dim)).astype(np.float32)
    inv_e, inv_i = (1.0 / pos_freqs).astype(np.float32), (1.0 / (np.float32(scale) * pos_freqs)).astype(np.float32)

    def corr(nrot):
        return np.float32(dim * np.log(np.float32(max_pos / (nrot * 2.0 * np.pi)))) / (2.0 * np.log(np.float32(theta)))
    low = float(max(0, int(np.floor(corr(beta_slow)))))
    high = float(min(dim - 1, int(np.ceil(corr(beta_fast)))))
    if abs(low - high) < 1e-6:
        high += 0.001
    ramp = np.clip((np.arange(dim // 2, dtype=np.float32) - low) / (high - low), 0, 1).astype(np.float32)
    mask = ((1.0 - ramp) * extrapolation_factor).astype(np.float32)
    inv_freq = (inv_i * (1.0 - mask) + inv_e * mask).astype(np.float32)
    t = np.arange(int(max_pos * scale), dtype=np.float32)
    freqs = np.outer(t, inv_freq).astype(np.float32)
    return (np.stack([np.cos(freqs), np.sin(freqs)], axis=-1).reshape(freqs.shape[0], -1) * np.float32(mscale)).astype(np.float32)


def qk_rmsnorm(qkv_bits, q_gamma_bits, k_gamma_bits, head_num, kv_head_num, head_dim, eps, q_bias_bits=None, k_bias_bits=None,
               is_bf16=False):
    x = np.ascontiguousarray(qkv_bits, np.uint16).copy()
    c = lambda a: None if a is None else _p(np.ascontiguousarray(a, np.uint16))
    qg, kg = np.ascontiguousarray(q_gamma_bits, np.uint16), np.ascontiguousarray(k_gamma_bits, np.uint16)
    qb = None if q_bias_bits is None else np.ascontiguousarray(q_bias_bits, np.uint16)
    kb = None if k_bias_bits is None else np.ascontiguousarray(k_bias_bits, np.uint16)
    lib().oracle_qk_rmsnorm(_p(x), _p(qg), _p(kg), None if qb is None else _p(qb), None if kb is None else _p(kb), int(is_bf16),
                            x.shape[0], head_num, kv_head_num, head_dim, ctypes.c_float(eps))
    return x
