"""Host-side mirror of the reference's decode-attention seam, over the C ABI.

  B200DecodeAttnOp   == the pybind class XQAAttnOp (support / prepare / update / update_kv_cache_offset / forward),
                        /root/reference/rtp_llm/models_py/bindings/cuda/XQAAttnOp.{h:13-28,cc:52-176}
  B200DecodeImpl     == an FMHAImplBase strategy (ctor (attn_configs, attn_inputs, parallelism_config), static support,
                        forward(qkv, kv_cache, layer_idx), prepare_cuda_graph(attn_inputs)),
                        /root/reference/rtp_llm/models_py/modules/factory/attention/fmha_impl_base.py:99-175 and
                        cuda_impl/xqa.py:61-153 (the structure mirrored here: rope+append op, then the paged attention op)

Inside the reference tree the impl is appended to DECODE_MHA_IMPS (attention/__init__.py:91-100); see INTEGRATION.md.
Config / input objects are duck-typed: any object with the reference's attribute names works (AttentionConfigs:
head_num, kv_head_num, size_per_head, kernel_tokens_per_block|tokens_per_block, max_seq_len, need_rope_kv_cache,
rope_config.base, q_scaling; PyAttentionInputs: sequence_lengths, kv_cache_kernel_block_id_device, is_prefill).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from ._lib import B200Error

try:  # inside the reference tree the real base class is used; standalone, a structural stand-in
    from rtp_llm.models_py.modules.factory.attention.fmha_impl_base import FMHAImplBase  # type: ignore
except Exception:  # noqa: BLE001
    class FMHAImplBase:  # same surface as fmha_impl_base.py:99-175
        accepts_fmha_config = False

        def support_cuda_graph(self) -> bool:
            return callable(getattr(self, "prepare_cuda_graph", None))

        @classmethod
        def support_parallelism_config(cls, parallelism_config) -> bool:
            return True


def _tokens_per_block(cfg) -> int:
    return int(getattr(cfg, "kernel_tokens_per_block", 0) or getattr(cfg, "tokens_per_block"))


@dataclass
class B200AttnParams:
    """== XQAParams (bindings/cuda/ops/CudaXqa.h:10-16): what prepare() hands to forward()."""
    kv_cache_offset: torch.Tensor          # [B,1,2,M] int32 page list
    sequence_lengths: torch.Tensor         # captured BY REFERENCE so graph replay can refresh it in place (xqa.py:78-80)
    batch_size: int
    max_seq_len: int
    workspace: torch.Tensor


class B200DecodeAttnOp:
    def __init__(self, attn_configs):
        self.cfg = attn_configs

    def support(self, attn_inputs) -> bool:
        c = self.cfg
        if getattr(attn_inputs, "is_prefill", False):
            return False
        group = c.head_num // max(c.kv_head_num, 1)
        ok = (c.size_per_head in (64, 128, 256) and c.head_num % max(c.kv_head_num, 1) == 0 and 1 <= group <= 16
              and _tokens_per_block(c) in (16, 32, 64, 128))
        kv_dtype = str(getattr(c, "kv_cache_dtype", "BASE"))
        return ok and "FP8" not in kv_dtype.upper() and torch.cuda.is_available() and \
            torch.cuda.get_device_capability()[0] == 10

    def prepare(self, attn_inputs) -> B200AttnParams:
        block_ids = attn_inputs.kv_cache_kernel_block_id_device
        if block_ids is None:
            raise B200Error("decode should have kv cache block id.")          # XQAAttnOp.cc:66-67
        B = int(attn_inputs.sequence_lengths.shape[0])
        if block_ids.shape[0] != B:
            raise B200Error(f"kv blocks batch size expected [{B}] but got [{block_ids.shape[0]}]")
        page_list = ops.convert_block_table(block_ids)
        max_seq_len = int(getattr(self.cfg, "max_seq_len", 0)) or int(block_ids.shape[1]) * _tokens_per_block(self.cfg)
        # sequence_lengths may equal max_seq_len (tokens already cached): the kernel covers positions 0..sequence_lengths
        # inclusive, so the bound is max_seq_len + 1 exactly as XQAAttnOp.cc:147-149 passes it, capped by the page table
        max_seq_len = min(max_seq_len + 1, int(block_ids.shape[1]) * _tokens_per_block(self.cfg))
        ws = ops.attn_workspace(B, self.cfg.head_num, self.cfg.kv_head_num, max_seq_len, block_ids.device)
        return B200AttnParams(page_list, attn_inputs.sequence_lengths, B, max_seq_len, ws)

    def update(self, params: B200AttnParams, attn_inputs) -> None:
        if params is None:
            raise B200Error("B200DecodeAttnOp::update received null params")
        self.update_kv_cache_offset(params.kv_cache_offset, attn_inputs.kv_cache_kernel_block_id_device)
        params.batch_size = int(attn_inputs.kv_cache_kernel_block_id_device.shape[0])
        params.sequence_lengths = attn_inputs.sequence_lengths

    def update_kv_cache_offset(self, kv_cache_offset: torch.Tensor, kv_cache_block_id_device: torch.Tensor) -> None:
        if kv_cache_offset.dim() != 4 or kv_cache_offset.shape[1] != 1 or kv_cache_offset.shape[2] != 2:
            raise B200Error("expects kv_cache_offset shape [batch, 1, 2, blocks]")          # XQAAttnOp.cc:31-32
        if kv_cache_offset.shape[0] != kv_cache_block_id_device.shape[0] or \
                kv_cache_offset.shape[3] != kv_cache_block_id_device.shape[1]:
            raise B200Error("shape mismatch: offset vs block table")
        ops.convert_block_table(kv_cache_block_id_device, out=kv_cache_offset)

    def forward(self, q: torch.Tensor, kv_cache, params: B200AttnParams) -> torch.Tensor:
        if kv_cache is None:
            raise B200Error("decode should have kv cache.")                                  # XQAAttnOp.cc:138
        base = kv_cache.kv_cache_base if hasattr(kv_cache, "kv_cache_base") else kv_cache
        seq = params.sequence_lengths
        if not seq.is_cuda and not seq.is_pinned():
            raise B200Error("sequence_lengths must be device-accessible (CUDA or pinned host memory)")
        q_scale = 1.0 / float(getattr(self.cfg, "q_scaling", 1.0) or 1.0)
        return ops.paged_decode_attn(q, base, params.kv_cache_offset, seq, params.max_seq_len, params.workspace,
                                     q_scale=q_scale)


class B200RopeKVCacheDecodeOp:
    """== FusedRopeKVCacheDecodeOp (rtp_llm/ops/fused_rope_kvcache_op.py:176-272), RopeStyle::Base / NeoX pairing."""

    def __init__(self, attn_configs):
        self.cfg = attn_configs

    def prepare(self, attn_inputs):
        return attn_inputs

    def forward(self, qkv: torch.Tensor, kv_cache, page_list: torch.Tensor, sequence_lengths: torch.Tensor):
        base = kv_cache.kv_cache_base if hasattr(kv_cache, "kv_cache_base") else kv_cache
        rope = getattr(self.cfg, "rope_config", None)
        rope_base = float(getattr(rope, "base", 10000.0)) if rope is not None else 10000.0
        return ops.rope_append(qkv, base, page_list, sequence_lengths, self.cfg.head_num, rope_base)


class B200DecodeImpl(FMHAImplBase):
    """Decode FMHA strategy; name contains no "XQA"/"TRT" so FMHAConfig toggles of other backends do not disable it
    (attn_factory.py:100-156)."""

    def __init__(self, attn_configs, attn_inputs, parallelism_config=None):
        self.need_rope_kv_cache = bool(getattr(attn_configs, "need_rope_kv_cache", True))
        self.fmha_impl = B200DecodeAttnOp(attn_configs)
        self.rope_kvcache_impl = B200RopeKVCacheDecodeOp(attn_configs)
        self.attn_inputs = attn_inputs
        self.fmha_params = self.fmha_impl.prepare(attn_inputs)
        self._captured_seq_lens = attn_inputs.sequence_lengths

    @staticmethod
    def support(attn_configs, attn_inputs) -> bool:
        return B200DecodeAttnOp(attn_configs).support(attn_inputs)

    def forward(self, qkv: torch.Tensor, kv_cache, layer_idx: int = 0) -> torch.Tensor:
        if self.need_rope_kv_cache:
            q = self.rope_kvcache_impl.forward(qkv, kv_cache, self.fmha_params.kv_cache_offset,
                                               self.fmha_params.sequence_lengths)
        else:
            q = qkv
        return self.fmha_impl.forward(q, kv_cache, self.fmha_params)

    def prepare_cuda_graph(self, attn_inputs) -> None:
        """Refresh per-step metadata IN PLACE (attention/common.py:88-138): page list and sequence lengths."""
        self.fmha_impl.update_kv_cache_offset(self.fmha_params.kv_cache_offset, attn_inputs.kv_cache_kernel_block_id_device)
        if attn_inputs.sequence_lengths is not self._captured_seq_lens:
            self._captured_seq_lens.copy_(attn_inputs.sequence_lengths, non_blocking=True)
