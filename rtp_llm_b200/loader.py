"""On-disk formats -> kernel layout: AutoGPTQ / AutoAWQ 4-bit group-wise checkpoints and plain FP16/BF16 weights for the
per-column INT8 path, read from safetensors and pushed through the device hooks exactly as the reference's loader does
(SURVEY 8 f-4).  What is mirrored, file:line in /root/reference:
  * tensor names   `<prefix>.qweight / .qzeros / .scales`                    model_loader/group_wise_quant_weight.py:32-36
  * which layers   qkv, o, w1 / w3 (merged into w13), w2                     group_wise_quant_weight.py:304-316
  * merge rules    q|k|v and gate|up concatenated along the OUTPUT axis (GPTQ qweight [K/8, N]: dim 1; AWQ qweight [K, N/8]: dim 1)
  * padding        the FFN inter size is padded to `align_size`: w2 along its INPUT rows (GPTQ qweight rows are packed 8:1 ->
                   align_size / 8; AWQ: align_size), zeros / scales by align_size / group_size rows; w1 / w3 along their
                   OUTPUT columns (group_wise_quant_weight.py:123-176, utils/model_weight.py:74-110 `pad`)
  * post-process   load_config.exported_device.preprocess_groupwise_weight_params(qweight, qzeros, scales, device, gptq, awq,
                   bits) (group_wise_quant_weight.py:392-425) -> here B200Impl (device.py); per-column INT8:
                   exported_device.apply_int8(kernel, device) (weight_only_quant_weight.py:94-105)
safetensors (the library) is plumbing for the file format, like torch is for device memory."""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional

import torch

from . import ops
from .device import B200Impl

QW, QZ, QS = ".qweight", ".qzeros", ".scales"


@dataclasses.dataclass
class QuantConfig:
    method: str              # "gptq" | "awq" | "int8" (per-column weight-only, quantised at load time) | "none"
    bits: int = 4
    group_size: int = 128

    @property
    def gptq(self) -> bool:
        return self.method == "gptq"

    @property
    def awq(self) -> bool:
        return self.method == "awq"


def pad_dim(t: torch.Tensor, align: int, dim: int, value: int = 0) -> torch.Tensor:
    """utils/model_weight.py:74-110: zero-pad `dim` up to a multiple of `align` (0 = no padding)."""
    if align <= 0:
        return t.contiguous()
    size = t.shape[dim]
    extra = (-size) % align
    if extra == 0:
        return t.contiguous()
    shape = list(t.shape)
    shape[dim] = extra
    return torch.cat([t, torch.full(shape, value, dtype=t.dtype, device=t.device)], dim=dim).contiguous()


class CheckpointReader:
    """Lazy view of one or more .safetensors files (AutoGPTQ / AutoAWQ / HF layout)."""

    def __init__(self, *paths: str):
        from safetensors import safe_open
        self._files = [safe_open(p, framework="pt", device="cpu") for p in paths]
        self._where = {k: f for f in self._files for k in f.keys()}

    def __contains__(self, name: str) -> bool:
        return name in self._where

    def get(self, name: str) -> torch.Tensor:
        if name not in self._where:
            raise KeyError(f"tensor {name!r} not found in the checkpoint")
        return self._where[name].get_tensor(name)


class B200Loader:
    """Builds the weights of one decoder layer in the kernel's layout. `names` maps the reference's logical weights to
    checkpoint prefixes, e.g. dict(q="model.layers.0.self_attn.q_proj", k=..., v=..., o=..., gate=..., up=..., down=...)."""

    def __init__(self, reader: CheckpointReader, quant: QuantConfig, device="cuda", act_dtype=torch.float16, align_size: int = 0,
                 fuse_silu: bool = True):
        self.r, self.q, self.device, self.act_dtype, self.align, self.fuse_silu = reader, quant, torch.device(device), act_dtype, align_size, fuse_silu
        self.impl = B200Impl(device, act_dtype)

    # ---- group-wise INT4 / INT8 (GPTQ / AWQ)
    def _triple(self, prefix: str):
        return self.r.get(prefix + QW), self.r.get(prefix + QZ), self.r.get(prefix + QS)

    def groupwise_tensors(self, prefixes, pad_in: int = 0, pad_out: int = 0):
        """Merge + pad + unpack on the host: the loader's UN-permuted tensors (q_packed uint8 [K, N/2] two's-complement
        nibbles, zeros_x_scales fp16 [K/g, N], scales fp16 [K/g, N]) -- what preprocess_groupwise_weight_params computes before
        its device-specific re-layout (device_impl.py:242-300)."""
        q = self.q
        if q.bits not in (4, 8):
            raise ValueError(f"group-wise checkpoints: bits must be 4 or 8, got {q.bits}")
        pad_div = 32 // q.bits                                        # group_wise_quant_weight.py:130
        qws, qzs, qss = zip(*[self._triple(p) for p in prefixes])
        if pad_out:                                                   # w1 / w3: the inter size is their OUTPUT axis
            qws = [pad_dim(w, pad_out if q.gptq else pad_out // pad_div, 1) for w in qws]
            qzs = [pad_dim(z, pad_out // pad_div, 1) for z in qzs]
            qss = [pad_dim(s, pad_out, 1) for s in qss]
        qw, qz, qs = torch.cat(qws, dim=1), torch.cat(qzs, dim=1), torch.cat(qss, dim=1)
        if pad_in:                                                    # w2: the inter size is its INPUT axis
            qw = pad_dim(qw, pad_in // pad_div if q.gptq else pad_in, 0)
            qz = pad_dim(qz, max(pad_in // q.group_size, 1), 0)
            qs = pad_dim(qs, max(pad_in // q.group_size, 1), 0)
        return self.impl.unpack_groupwise(qw, qz, qs.to(torch.float16), q.gptq, q.awq, q.bits)

    def _groupwise(self, prefixes, pad_in: int = 0, pad_out: int = 0, gate_up_inter: int = 0) -> ops.PackedWeight:
        q_packed, zs, scales = self.groupwise_tensors(prefixes, pad_in, pad_out)
        q_packed, zs, scales = q_packed.to(self.device), zs.to(self.act_dtype).to(self.device), scales.to(self.act_dtype).to(self.device)
        if gate_up_inter:
            q_packed = ops.interleave_gate_up(q_packed, gate_up_inter, packed_int4=self.q.bits == 4)
            zs, scales = ops.interleave_gate_up(zs, gate_up_inter), ops.interleave_gate_up(scales, gate_up_inter)
        pack = ops.pack_w4 if self.q.bits == 4 else ops.pack_w8g     # 8-bit: q_s int8 [K, N], one byte per weight
        return pack(q_packed.contiguous(), scales.contiguous(), zs.contiguous(), self.q.group_size)

    # ---- per-column INT8 (quantised at load time from FP16/BF16 weights) and plain FP16
    def _dense(self, prefixes, pad_in: int = 0, pad_out: int = 0, gate_up_inter: int = 0) -> ops.PackedWeight:
        ws = [self.r.get(p + ".weight").t().contiguous() for p in prefixes]          # HF stores [N, K]; the reference works on [K, N]
        if pad_out:
            ws = [pad_dim(w, pad_out, 1) for w in ws]
        w = torch.cat(ws, dim=1)
        if pad_in:
            w = pad_dim(w, pad_in, 0)
        if self.q.method == "int8":
            qv, scale = self.impl.symmetric_quantize_last_axis_of_batched_matrix(w.float())
            qv, scale = qv.to(self.device), scale.to(self.act_dtype).to(self.device)
            if gate_up_inter:
                qv, scale = ops.interleave_gate_up(qv, gate_up_inter), ops.interleave_gate_up(scale, gate_up_inter)
            return ops.pack_w8(qv.contiguous(), scale.contiguous())
        w = w.to(self.act_dtype).to(self.device)
        if gate_up_inter:
            w = ops.interleave_gate_up(w, gate_up_inter)
        return ops.pack_f16(w)

    def _load(self, prefixes, **kw) -> ops.PackedWeight:
        return self._groupwise(prefixes, **kw) if self.q.method in ("gptq", "awq") else self._dense(prefixes, **kw)

    def layer(self, names: Dict[str, str], inter: Optional[int] = None) -> Dict[str, ops.PackedWeight]:
        """qkv / o / w13 / w2 of one decoder layer. `inter` = the checkpoint's FFN inter size (needed for padding + fusion)."""
        out = {"qkv": self._load([names["q"], names["k"], names["v"]]), "o": self._load([names["o"]])}
        pad = self.align
        inter_p = inter if inter is None or not pad else (inter + pad - 1) // pad * pad
        fuse = self.fuse_silu and inter_p is not None and inter_p % 64 == 0
        out["w13"] = self._load([names["gate"], names["up"]], pad_out=pad, gate_up_inter=inter_p if fuse else 0)
        out["w2"] = self._load([names["down"]], pad_in=pad)
        out["w13_fused_silu"] = fuse
        return out
