"""Torch-tensor face of the C ABI. PyTorch is plumbing here (device memory + streams); all compute is in
libb200_decode.so. Every function launches on the current CUDA stream and is CUDA-graph capturable."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import B200_FMT_F16, B200_FMT_INT4, B200_FMT_INT8, B200_FMT_INT8G, B200Error, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _is_bf16(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float16:
        return 0
    raise B200Error(f"unsupported activation dtype {t.dtype} (fp16 / bf16 only)")


def _cuda_contig(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not (t.is_cuda and t.is_contiguous()):
            raise B200Error("expected contiguous CUDA tensors")


def device_check(device: int = 0) -> None:
    check(_lib.load().b200_device_check(device), "b200_device_check")


def set_pdl(enable: bool) -> None:
    """Programmatic dependent launch for the whole decode chain (see include/b200_decode_ops.h)."""
    check(_lib.load().b200_set_pdl(1 if enable else 0), "b200_set_pdl")


def launch_count() -> int:
    return int(_lib.load().b200_launch_count())


# ------------------------------------------------------------------------------------------------ indexing
def convert_block_table(block_ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,M] int32 -> [B,1,2,M] int32 page list (XQAAttnOp::prepare, XQAAttnOp.cc:77-92)."""
    _cuda_contig(block_ids, out)
    B, M = block_ids.shape
    if out is None:
        out = torch.empty((B, 1, 2, M), dtype=torch.int32, device=block_ids.device)
    check(_lib.load().b200_convert_block_table(_p(out), _p(block_ids), B, M, _stream()), "b200_convert_block_table")
    return out


def paged_attn_plan(sequence_lengths: torch.Tensor, block_ids: Optional[torch.Tensor], tokens_per_block: int,
                    input_lengths: Optional[torch.Tensor] = None, prefix_lengths: Optional[torch.Tensor] = None):
    ref = sequence_lengths if prefix_lengths is None else prefix_lengths
    B = ref.shape[0]
    dev = ref.device
    M = block_ids.shape[1] if block_ids is not None else 0
    ntok = B if prefix_lengths is None else int(input_lengths.sum().item())
    last = torch.zeros(B, dtype=torch.int32, device=dev)
    indptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    indice = torch.zeros(max(B * M, 1), dtype=torch.int32, device=dev)
    bidx = torch.zeros(max(ntok, 1), dtype=torch.int32, device=dev)
    pos = torch.zeros(max(ntok, 1), dtype=torch.int32, device=dev)
    check(_lib.load().b200_paged_attn_plan(_p(input_lengths), _p(sequence_lengths), _p(prefix_lengths), _p(block_ids), B,
                                           M, tokens_per_block, _p(last), _p(indptr), _p(indice), _p(bidx), _p(pos),
                                           _stream()), "b200_paged_attn_plan")
    return dict(last_page_len=last, page_indptr=indptr, page_indice=indice, batch_indice=bidx[:ntok], positions=pos[:ntok])


# ------------------------------------------------------------------------------------------------ attention
def attn_workspace(batch: int, head_num: int, kv_head_num: int, max_seq_len: int, device) -> torch.Tensor:
    n = _lib.load().b200_paged_decode_attn_workspace_bytes(batch, head_num, kv_head_num, max_seq_len)
    return torch.zeros(max(int(n), 256), dtype=torch.uint8, device=device)


def paged_decode_attn(q: torch.Tensor, kv_cache_base: torch.Tensor, page_list: torch.Tensor,
                      sequence_lengths: torch.Tensor, max_seq_len: int, workspace: torch.Tensor,
                      out: Optional[torch.Tensor] = None, q_scale: float = 1.0) -> torch.Tensor:
    """q [B,Hq,D] (or [B,Hq*D]); kv_cache_base [P,2,Hkv,T,D]; page_list [B,1,2,M]; sequence_lengths [B] int32 = tokens
    already cached; max_seq_len = host bound on sequence_lengths+1. Returns [B, Hq*D]."""
    _cuda_contig(q, kv_cache_base, page_list, workspace, out)
    P, two, Hkv, T, D = kv_cache_base.shape
    B = q.shape[0]
    Hq = q.numel() // (B * D)
    if out is None:
        out = torch.empty((B, Hq * D), dtype=q.dtype, device=q.device)
    check(_lib.load().b200_paged_decode_attn(_p(q), _is_bf16(q), _p(out), Hq, Hkv, D, B, page_list.shape[-1], max_seq_len,
                                             T, _p(kv_cache_base), _p(page_list), _p(sequence_lengths), q_scale,
                                             _p(workspace), workspace.numel(), _stream()), "b200_paged_decode_attn")
    return out


def paged_decode_attn_multi(q: torch.Tensor, kv_cache_base: torch.Tensor, page_list: torch.Tensor, sequence_lengths: torch.Tensor,
                            max_seq_len: int, workspace: torch.Tensor, out: Optional[torch.Tensor] = None, q_scale: float = 1.0) -> torch.Tensor:
    """q [B, q_len, Hq, D] (q_len speculative tokens per sequence, K/V already appended); sequence_lengths [B] = tokens cached
    before them; query j attends to positions 0 .. sequence_lengths[b] + j. Returns [B, q_len, Hq*D]."""
    _cuda_contig(q, kv_cache_base, page_list, workspace, out)
    P, two, Hkv, T, D = kv_cache_base.shape
    B, q_len, Hq = q.shape[0], q.shape[1], q.shape[2]
    if out is None:
        out = torch.empty((B, q_len, Hq * D), dtype=q.dtype, device=q.device)
    check(_lib.load().b200_paged_decode_attn_multi(_p(q), _is_bf16(q), _p(out), Hq, Hkv, D, B, q_len, page_list.shape[-1], max_seq_len,
                                                   T, _p(kv_cache_base), _p(page_list), _p(sequence_lengths), q_scale,
                                                   _p(workspace), workspace.numel(), _stream()), "b200_paged_decode_attn_multi")
    return out


def paged_decode_attn_rope(qkv: torch.Tensor, kv_cache_base: torch.Tensor, page_list: torch.Tensor, sequence_lengths: torch.Tensor,
                           head_num: int, max_seq_len: int, rope_base: float, workspace: torch.Tensor,
                           out: Optional[torch.Tensor] = None, q_scale: float = 1.0) -> torch.Tensor:
    """RoPE + K/V append + paged decode attention in ONE launch: qkv [B, (Hq+2Hkv)*D] un-rotated (the qkv GEMM output);
    kv_cache_base receives the new token's K (rotated) and V. Bit-identical to rope_append() followed by paged_decode_attn()."""
    _cuda_contig(qkv, kv_cache_base, page_list, workspace, out)
    P, two, Hkv, T, D = kv_cache_base.shape
    B = qkv.shape[0]
    if out is None:
        out = torch.empty((B, head_num * D), dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().b200_paged_decode_attn_rope(_p(qkv), _is_bf16(qkv), _p(out), head_num, Hkv, D, B, page_list.shape[-1], max_seq_len,
                                                  T, _p(kv_cache_base), _p(page_list), _p(sequence_lengths), q_scale, rope_base,
                                                  _p(workspace), workspace.numel(), _stream()), "b200_paged_decode_attn_rope")
    return out


# ------------------------------------------------------------------------------------------------ weight-only GEMM
_TRAILER_BYTES = 256
_TRAILER_MAGIC = 0x42323030574F4731        # "B200WOG1"


class PackedWeight:
    """A weight in the layout b200_wo_gemm consumes (+ what the epilogue needs).

    Quantised blobs are SELF-DESCRIBING: the uint8 tensor is [blob bytes | 256-byte trailer] with a magic number, the format and
    (K, N, group). The reference loader hands kernels around as plain tensors and copies them freely
    (device_impl.py:296-298 `.contiguous().to(device)`, state-dict round trips): a Python attribute on the tensor would be
    lost there, the trailer travels with the bytes. PackedWeight.from_tensor() rebuilds the description from any copy."""

    def __init__(self, fmt: int, K: int, N: int, data: torch.Tensor, col_scale: Optional[torch.Tensor] = None):
        self.fmt, self.K, self.N, self.data, self.col_scale = fmt, K, N, data, col_scale

    @staticmethod
    def _write_trailer(data: torch.Tensor, fmt: int, K: int, N: int, group: int) -> None:
        t = torch.tensor([_TRAILER_MAGIC, fmt, K, N, group], dtype=torch.int64)
        data[-_TRAILER_BYTES:-_TRAILER_BYTES + 40] = t.view(torch.uint8).to(data.device)

    @staticmethod
    def from_tensor(data: torch.Tensor, col_scale: Optional[torch.Tensor] = None) -> "PackedWeight":
        if data.dtype != torch.uint8 or data.dim() != 1 or data.numel() <= _TRAILER_BYTES:
            raise B200Error("not a b200 weight blob (expected a 1-D uint8 tensor with a trailer)")
        t = data[-_TRAILER_BYTES:-_TRAILER_BYTES + 40].cpu().view(torch.int64).tolist()
        if t[0] != _TRAILER_MAGIC:
            raise B200Error("not a b200 weight blob (bad magic): quantised weights must be loaded through B200Impl")
        fmt, K, N = int(t[1]), int(t[2]), int(t[3])
        n = int(_lib.load().b200_wo_gemm_packed_bytes(fmt, K, N))
        if n + _TRAILER_BYTES != data.numel():
            raise B200Error(f"b200 weight blob has {data.numel()} bytes, expected {n + _TRAILER_BYTES} for K={K} N={N}")
        return PackedWeight(fmt, K, N, data if data.is_contiguous() else data.contiguous(), col_scale)


def pack_w4(q_packed: torch.Tensor, scales: torch.Tensor, zeros_x_scales: torch.Tensor, group: int = 128) -> PackedWeight:
    """q_packed uint8/int8 [K,N/2], scales / zeros_x_scales [K/g,N] in the activation dtype (CUDA)."""
    _cuda_contig(q_packed, scales, zeros_x_scales)
    K, N = q_packed.shape[0], q_packed.shape[1] * 2
    n = _lib.load().b200_wo_gemm_packed_bytes(B200_FMT_INT4, K, N)
    if n == 0:
        raise B200Error(f"pack_w4: unsupported shape K={K} N={N}")
    blob = torch.empty(int(n) + _TRAILER_BYTES, dtype=torch.uint8, device=q_packed.device)
    check(_lib.load().b200_pack_w4(_p(q_packed), _p(scales), _p(zeros_x_scales), K, N, group, _p(blob), _stream()),
          "b200_pack_w4")
    PackedWeight._write_trailer(blob, B200_FMT_INT4, K, N, group)
    return PackedWeight(B200_FMT_INT4, K, N, blob)


def pack_w8(q: torch.Tensor, col_scale: torch.Tensor) -> PackedWeight:
    _cuda_contig(q, col_scale)
    K, N = q.shape
    n = _lib.load().b200_wo_gemm_packed_bytes(B200_FMT_INT8, K, N)
    if n == 0:
        raise B200Error(f"pack_w8: unsupported shape K={K} N={N}")
    blob = torch.empty(int(n) + _TRAILER_BYTES, dtype=torch.uint8, device=q.device)
    check(_lib.load().b200_pack_w8(_p(q), K, N, _p(blob), _stream()), "b200_pack_w8")
    PackedWeight._write_trailer(blob, B200_FMT_INT8, K, N, 0)
    return PackedWeight(B200_FMT_INT8, K, N, blob, col_scale)


def pack_w8g(q: torch.Tensor, scales: torch.Tensor, zeros_x_scales: torch.Tensor, group: int = 128) -> PackedWeight:
    """8-bit group-wise (GPTQ/AWQ W8): q int8 [K,N] (q_s = q_u - 128), scales / zeros_x_scales [K/g,N] in the activation dtype."""
    _cuda_contig(q, scales, zeros_x_scales)
    K, N = q.shape
    n = _lib.load().b200_wo_gemm_packed_bytes(B200_FMT_INT8G, K, N)
    if n == 0:
        raise B200Error(f"pack_w8g: unsupported shape K={K} N={N}")
    blob = torch.empty(int(n) + _TRAILER_BYTES, dtype=torch.uint8, device=q.device)
    check(_lib.load().b200_pack_w8g(_p(q), _p(scales), _p(zeros_x_scales), K, N, group, _p(blob), _stream()), "b200_pack_w8g")
    PackedWeight._write_trailer(blob, B200_FMT_INT8G, K, N, group)
    return PackedWeight(B200_FMT_INT8G, K, N, blob)


def pack_f16(w_kn: torch.Tensor) -> PackedWeight:
    """Reference stores W [K,N]; the kernel wants K contiguous -> one transpose at load time."""
    K, N = w_kn.shape
    return PackedWeight(B200_FMT_F16, K, N, w_kn.t().contiguous())


def gemm_workspace(max_batch: int, shapes, device) -> torch.Tensor:
    """One zero-filled scratch buffer big enough for every (K, N) in `shapes` (used serially on one stream)."""
    n = max(int(_lib.load().b200_wo_gemm_workspace_bytes(max_batch, N, K)) for (K, N) in shapes)
    return torch.zeros(max(n, 256), dtype=torch.uint8, device=device)


def gate_up_order(inter: int, device=None) -> torch.Tensor:
    """Column order that puts gate j and up j into ADJACENT rows of a 128-feature tile (row 2i = gate tile*64+i, row 2i+1 =
    up tile*64+i): the layout the fused SiLU*mul epilogue (B200_GEMM_SILU_MUL) expects of a [K, 2*inter] gate|up weight --
    the two values of a pair then sit in neighbouring lanes of one warp and are exchanged with a shuffle."""
    assert inter % 64 == 0
    j = torch.arange(inter, device=device)
    return torch.stack([j, j + inter], dim=1).reshape(-1)


def interleave_gate_up(w: torch.Tensor, inter: int, packed_int4: bool = False) -> torch.Tensor:
    """Apply gate_up_order to the column axis of a reference-layout tensor ([K, 2I], or uint8 [K, I] packed int4 -- low
    nibble = even column --, or scales [G, 2I] / [2I])."""
    order = gate_up_order(inter, w.device)
    if packed_int4:
        full = torch.stack([w & 0xF, w >> 4], dim=-1).reshape(w.shape[0], -1)       # one nibble per column [K, 2I]
        full = full.index_select(-1, order)
        return (full[:, 0::2] | (full[:, 1::2] << 4)).contiguous()
    return w.index_select(-1, order).contiguous()


def wo_gemm(x: torch.Tensor, w: PackedWeight, workspace: torch.Tensor, bias: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, pdl: bool = False, silu_mul: bool = False) -> torch.Tensor:
    _cuda_contig(x, w.data, workspace, bias, out)
    B, K = x.shape
    if K != w.K:
        raise B200Error(f"wo_gemm: x has K={K}, weight has K={w.K}")
    if out is None:
        out = torch.empty((B, w.N // 2 if silu_mul else w.N), dtype=x.dtype, device=x.device)
    check(_lib.load().b200_wo_gemm(w.fmt, _is_bf16(x), _p(x), B, K, w.N, _p(w.data), _p(w.col_scale), _p(bias), _p(out),
                                   _p(workspace), workspace.numel(),
                                   (_lib.B200_GEMM_PDL if pdl else 0) | (_lib.B200_GEMM_SILU_MUL if silu_mul else 0), _stream()),
          "b200_wo_gemm")
    return out


# ------------------------------------------------------------------------------------------------ glue ops
def add_rmsnorm(x, residual, gamma, eps, out=None):
    _cuda_contig(x, residual, gamma, out)
    rows, hidden = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().b200_add_rmsnorm(_p(x), _p(residual), _p(gamma), _p(out), _is_bf16(x), rows, hidden, eps, _stream()),
          "b200_add_rmsnorm")
    return out


def qk_rmsnorm(qkv, q_gamma, k_gamma, head_num, kv_head_num, head_dim, eps, q_bias=None, k_bias=None):
    """In-place per-head RMSNorm of the q and k heads of qkv [rows, (Hq+2Hkv)*D] (fused_qk_rmsnorm.cu)."""
    _cuda_contig(qkv, q_gamma, k_gamma, q_bias, k_bias)
    check(_lib.load().b200_qk_rmsnorm(_p(qkv), _p(q_gamma), _p(k_gamma), _p(q_bias), _p(k_bias), _is_bf16(qkv), qkv.shape[0], head_num,
                                      kv_head_num, head_dim, eps, _stream()), "b200_qk_rmsnorm")
    return qkv


def silu_and_mul(gate_up, out=None):
    _cuda_contig(gate_up, out)
    rows, two_inter = gate_up.shape
    if out is None:
        out = torch.empty((rows, two_inter // 2), dtype=gate_up.dtype, device=gate_up.device)
    check(_lib.load().b200_silu_and_mul(_p(gate_up), _p(out), _is_bf16(gate_up), rows, two_inter // 2, _stream()),
          "b200_silu_and_mul")
    return out


def rope_append(qkv, kv_cache_base, page_list, sequence_lengths, head_num, rope_base, q_out=None):
    _cuda_contig(qkv, kv_cache_base, page_list, sequence_lengths, q_out)
    P, two, Hkv, T, D = kv_cache_base.shape
    B = qkv.shape[0]
    if q_out is None:
        q_out = torch.empty((B, head_num * D), dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().b200_rope_append(_p(qkv), _p(q_out), _p(kv_cache_base), _p(page_list), _p(sequence_lengths),
                                       _is_bf16(qkv), B, head_num, Hkv, D, page_list.shape[-1], T, rope_base, _stream()),
          "b200_rope_append")
    return q_out


def rope_append_ex(qkv, kv_cache_base, page_list, sequence_lengths, head_num, rope_cfg: "_lib.RopeConfig", q_out=None, bias=None,
                   position_ids=None, cos_sin_cache=None, use_logn_attn=False):
    """FusedRopeKVCacheDecodeOp.forward with the whole rope contract (styles, position_ids override, bias, cos/sin cache,
    logn): see b200_rope_append_ex in include/b200_decode_ops.h. cos_sin_cache: fp32 [positions, dim] interleaved (cos, sin)."""
    import ctypes
    _cuda_contig(qkv, kv_cache_base, page_list, sequence_lengths, q_out, bias, position_ids, cos_sin_cache)
    P, two, Hkv, T, D = kv_cache_base.shape
    B = qkv.shape[0]
    if q_out is None:
        q_out = torch.empty((B, head_num * D), dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().b200_rope_append_ex(_p(qkv), _p(bias), _p(q_out), _p(kv_cache_base), _p(page_list), _p(sequence_lengths),
                                          _p(position_ids), _p(cos_sin_cache), cos_sin_cache.shape[0] if cos_sin_cache is not None else 0,
                                          ctypes.cast(ctypes.pointer(rope_cfg), ctypes.c_void_p), 1 if use_logn_attn else 0,
                                          _is_bf16(qkv), B, head_num, Hkv, D, page_list.shape[-1], T, _stream()), "b200_rope_append_ex")
    return q_out


def embedding(ids, table, out=None):
    _cuda_contig(ids, table, out)
    rows, hidden = ids.shape[0], table.shape[1]
    if out is None:
        out = torch.empty((rows, hidden), dtype=table.dtype, device=table.device)
    check(_lib.load().b200_embedding(_p(ids), _p(table), _p(out), _is_bf16(table), rows, hidden, _stream()), "b200_embedding")
    return out


def argmax(logits, out=None):
    _cuda_contig(logits, out)
    rows, vocab = logits.shape
    dt = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}[logits.dtype]
    if out is None:
        out = torch.empty(rows, dtype=torch.int32, device=logits.device)
    check(_lib.load().b200_argmax(_p(logits), dt, rows, vocab, _p(out), _stream()), "b200_argmax")
    return out


def sample(logits: torch.Tensor, top_k: torch.Tensor, top_p: torch.Tensor, uniform: torch.Tensor,
           temperature: Optional[torch.Tensor] = None, history: Optional[torch.Tensor] = None,
           hist_len: Optional[torch.Tensor] = None, repetition: Optional[torch.Tensor] = None,
           presence: Optional[torch.Tensor] = None, frequency: Optional[torch.Tensor] = None,
           process: Optional[torch.Tensor] = None, count_ws: Optional[torch.Tensor] = None, want_probs: bool = False):
    """sampleGreedy's CUDA path (CudaSampleOp.cc:423-463) on fp32 logits [B, V], IN PLACE (logits become probabilities).
    Returns (tokens int32 [B], token_prob fp32 [B], renormalised probs [B, V] or None)."""
    _cuda_contig(logits, top_k, top_p, uniform, temperature, history, hist_len, repetition, presence, frequency, process, count_ws)
    if logits.dtype != torch.float32:
        raise B200Error("sample: logits must be fp32 (the reference sampler runs on fp32 logits)")
    B, V = logits.shape
    if (repetition is not None or presence is not None or frequency is not None) and count_ws is None:
        count_ws = torch.zeros(B, V, dtype=torch.int32, device=logits.device)
    tok = torch.empty(B, dtype=torch.int32, device=logits.device)
    tprob = torch.empty(B, dtype=torch.float32, device=logits.device)
    probs = torch.empty_like(logits) if want_probs else None
    check(_lib.load().b200_sample(_p(logits), B, V, _p(history), _p(hist_len), history.shape[1] if history is not None else 0,
                                  _p(count_ws), _p(temperature), _p(repetition), _p(presence), _p(frequency), _p(top_k), _p(top_p),
                                  _p(uniform), _p(process), _p(tok), _p(tprob), _p(probs), _stream()), "b200_sample")
    return tok, tprob, probs


# ------------------------------------------------------------------------------------------------ decode programs
class Program:
    """Recorded sequence of op calls (b200_program_*): `with prog.record(): <ops...>` then `prog.launch()` replays them
    with the GEMMs / norms / rope between two attention calls fused into one persistent kernel. Pointers are frozen at
    record time (like a CUDA-graph capture); launch() is itself graph-capturable."""

    def __init__(self):
        import ctypes
        self._h = ctypes.c_void_p()
        check(_lib.load().b200_program_create(ctypes.byref(self._h)), "b200_program_create")
        self._keep = []

    class _Rec:
        def __init__(self, prog):
            self.prog = prog

        def __enter__(self):
            check(_lib.load().b200_program_begin(self.prog._h), "b200_program_begin")
            return self.prog

        def __exit__(self, et, ev, tb):
            rc = _lib.load().b200_program_end(self.prog._h)
            if et is None:
                check(rc, "b200_program_end")
            return False

    def record(self):
        return Program._Rec(self)

    def launch(self) -> None:
        check(_lib.load().b200_program_launch(self._h, _stream()), "b200_program_launch")

    @property
    def num_ops(self) -> int:
        return int(_lib.load().b200_program_num_ops(self._h))

    @property
    def num_launches(self) -> int:
        return int(_lib.load().b200_program_num_launches(self._h))

    def set_trace(self, buf: Optional[torch.Tensor]) -> None:
        self._keep.append(buf)
        check(_lib.load().b200_program_set_trace(self._h, _p(buf)), "b200_program_set_trace")

    def __del__(self):
        try:
            if self._h:
                _lib.load().b200_program_destroy(self._h)
                self._h = None
        except Exception:  # noqa: BLE001
            pass
