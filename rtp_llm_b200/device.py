"""Host-side mirror of the loader's device hooks: DeviceBase.{apply_int8, preprocess_groupwise_weight_params,
preprocess_weights_for_mixed_gemm} (/root/reference/rtp_llm/device/device_base.py:56-90), instance stored in
LoadConfig.exported_device (model_loader/loader.py:79).

The unpack / quantise arithmetic restates device_impl.py:148-300 in torch (bit-exact against the reference's code on the
golden fixtures, tests/test_host_logic.py); only the final re-layout differs: instead of the FT sm80 interleave
(device_impl.py:392-479) weights go into the TMA-friendly blobs of include/b200_decode_ops.h."""
from __future__ import annotations

import torch

from . import ops


def unpack_int32_into_int16(w_packed: torch.Tensor, int8: bool = False) -> torch.Tensor:
    """device_impl.py:147-161: int32 words -> bytes (8-bit checkpoints) or nibbles (low nibble first) along the last axis."""
    b = w_packed.contiguous().view(torch.uint8)
    if int8:
        return b.to(torch.int16)
    out = torch.empty(b.shape[0], b.shape[1] * 2, dtype=torch.int16, device=b.device)
    out[:, ::2] = (b % 16).to(torch.int16)
    out[:, 1::2] = (b // 16).to(torch.int16)
    return out


def reverse_awq_order(t: torch.Tensor) -> torch.Tensor:
    """device_impl.py:163-171."""
    return t.reshape(-1, 2, 4).transpose(2, 1).reshape(t.shape)


def pack_int8_tensor_to_packed_int4(t: torch.Tensor) -> torch.Tensor:
    """device_impl.py:204-209: two's-complement nibbles, low nibble = even column."""
    u = (t.to(torch.int16) & 0xF).to(torch.uint8)
    return (u[:, 1::2] * 16 + u[:, ::2]).contiguous()


class B200Impl:
    """Drop-in for CudaImpl's weight hooks. The returned `kernel` tensors are self-describing blobs (ops.PackedWeight: a
    trailer inside the tensor names format / K / N), so they survive the loader's `.contiguous().to(device)`, clones and
    state-dict round trips; B200WeightOnlyLinear rebuilds the description with PackedWeight.from_tensor()."""

    def __init__(self, device="cuda", act_dtype=torch.float16):
        self.device, self.act_dtype = torch.device(device), act_dtype

    # -- device_impl.py:183-202 / 215-222
    def symmetric_quantize_last_axis_of_batched_matrix(self, weight: torch.Tensor, quant_mode=torch.int8):
        amax = torch.clamp(weight.abs().max(dim=0)[0], min=1e-8)
        scale = amax / 128.0
        q = torch.clamp((weight / scale).round(), -128, 127).char()
        return q, scale

    def apply_int8(self, tensor: torch.Tensor, device: str = "cuda"):
        shape = tensor.shape
        q, scale = self.symmetric_quantize_last_axis_of_batched_matrix(tensor.reshape(shape[0], -1).float())
        packed = ops.pack_w8(q.to(self.device).contiguous(), scale.to(self.act_dtype).to(self.device))
        return packed.data, scale.to(self.device)

    # -- device_impl.py:242-300
    def unpack_groupwise(self, qweight_int32, qzeros_int32, scales_fp16, gptq: bool, awq: bool, weight_bits: int = 4):
        """Returns the loader's UN-permuted tensors: 4-bit (q_packed uint8 [K,N/2], zeros_x_scales fp16, scales fp16);
        8-bit (device_impl.py:256-258: zero shift 128, no nibble packing) (q_s int8 [K,N], zeros_x_scales, scales)."""
        if weight_bits not in (4, 8):
            raise ValueError(f"group-wise weight_bits must be 4 or 8, got {weight_bits}")
        is_int8 = weight_bits == 8
        shift = 128 if is_int8 else 8
        qweight = qweight_int32.reshape(qweight_int32.shape[0], -1)
        qzeros = qzeros_int32.reshape(qzeros_int32.shape[0], -1)
        scales = scales_fp16.reshape(scales_fp16.shape[0], -1)
        if awq:
            q = reverse_awq_order(unpack_int32_into_int16(qweight, is_int8) - shift)
        elif gptq:
            q = (unpack_int32_into_int16(qweight.T.contiguous(), is_int8).T.contiguous() - shift)
        else:
            raise ValueError("need gptq or awq")
        q = q.to(torch.int8)
        q_out = q.contiguous() if is_int8 else pack_int8_tensor_to_packed_int4(q)
        z = unpack_int32_into_int16(qzeros, is_int8)
        if awq:
            z = reverse_awq_order(z)
        zeros_x_scales = ((-z + shift - (1 if gptq else 0)) * scales).half()
        return q_out, zeros_x_scales, scales

    def preprocess_groupwise_weight_params(self, qweight_int32, qzeros_int32, scales_fp16, device: str, gptq: bool,
                                           awq: bool, weight_bits: int):
        q_packed, zs, scales = self.unpack_groupwise(qweight_int32, qzeros_int32, scales_fp16, gptq, awq, weight_bits)
        kernel = self.preprocess_weights_for_mixed_gemm(q_packed, torch.int8 if weight_bits == 8 else torch.quint4x2, scales=scales,
                                                        zeros_x_scales=zs)
        return kernel, zs.to(self.device), scales.to(self.device)

    # -- replaces device_impl.py:392-479
    def preprocess_weights_for_mixed_gemm(self, tensor: torch.Tensor, quant_mode, arch: str = "", scales=None,
                                          zeros_x_scales=None):
        if quant_mode == torch.int8 and zeros_x_scales is None:
            raise ValueError("INT8 per-column weights are packed by apply_int8 (the scale rides along)")
        assert scales is not None and zeros_x_scales is not None, "the b200 group-wise blobs carry scales and zero*scale"
        if quant_mode == torch.int8:        # 8-bit group-wise
            return ops.pack_w8g(tensor.to(self.device).contiguous(), scales.to(self.act_dtype).to(self.device).contiguous(),
                                zeros_x_scales.to(self.act_dtype).to(self.device).contiguous()).data
        packed = ops.pack_w4(tensor.to(self.device).contiguous(), scales.to(self.act_dtype).to(self.device).contiguous(),
                             zeros_x_scales.to(self.act_dtype).to(self.device).contiguous())
        return packed.data
