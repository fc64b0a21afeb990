"""Loader (rtp_llm_b200/loader.py): synthetic AutoGPTQ / AutoAWQ safetensors checkpoints -> the loader's un-permuted tensors,
checked against the dense weight the checkpoint MEANS (independent definition of the two formats, tests/ckpt_util.py);
group padding of the FFN inter size as group_wise_quant_weight.py:123-176 does it. CPU only: the device re-layout and the GEMM
are covered by tests/test_gpu_loader.py."""
import numpy as np
import pytest
import torch

from rtp_llm_b200.loader import B200Loader, CheckpointReader, QuantConfig, pad_dim
from tests import ckpt_util


def _dense_from_unpacked(q_packed, zs, scales, group=128):
    if q_packed.dtype == torch.int8:                                         # 8-bit group-wise: q_s, one byte per weight
        return (q_packed.numpy().astype(np.float32) * np.repeat(scales.float().numpy(), group, 0)
                + np.repeat(zs.float().numpy(), group, 0))
    b = q_packed.numpy().astype(np.uint8)
    lo, hi = (b & 0xF).astype(np.int16), (b >> 4).astype(np.int16)
    q = np.empty((b.shape[0], b.shape[1] * 2), np.int16)
    q[:, 0::2], q[:, 1::2] = lo, hi
    q = np.where(q >= 8, q - 16, q).astype(np.float32)                       # two's-complement nibbles
    return q * np.repeat(scales.float().numpy(), group, 0) + np.repeat(zs.float().numpy(), group, 0)


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("method", ["gptq", "awq"])
def test_groupwise_checkpoint_unpacks_to_its_dense_meaning(tmp_path, method, bits):
    rng = np.random.default_rng(5)
    path = str(tmp_path / "layer.safetensors")
    hidden, heads, kvh, D, inter = 256, 4, 2, 64, 384
    names, dense = ckpt_util.write_checkpoint(path, rng, method, hidden, heads, kvh, D, inter, bits=bits)
    ld = B200Loader(CheckpointReader(path), QuantConfig(method, bits=bits), device="cpu")
    # qkv: q | k | v merged along the output axis
    got = _dense_from_unpacked(*ld.groupwise_tensors([names["q"], names["k"], names["v"]]))
    np.testing.assert_allclose(got, np.concatenate([dense["q"], dense["k"], dense["v"]], 1), atol=2e-3, rtol=2e-3)
    np.testing.assert_allclose(_dense_from_unpacked(*ld.groupwise_tensors([names["o"]])), dense["o"], atol=2e-3, rtol=2e-3)
    # FFN with the inter size padded from 384 to 512: w13 gains zero output columns per half, w2 zero input rows
    w13 = _dense_from_unpacked(*ld.groupwise_tensors([names["gate"], names["up"]], pad_out=512))
    assert w13.shape == (hidden, 1024)
    np.testing.assert_allclose(w13[:, :384], dense["gate"], atol=2e-3, rtol=2e-3)
    np.testing.assert_allclose(w13[:, 512:896], dense["up"], atol=2e-3, rtol=2e-3)
    assert np.all(w13[:, 384:512] == 0) and np.all(w13[:, 896:] == 0)
    w2 = _dense_from_unpacked(*ld.groupwise_tensors([names["down"]], pad_in=512))
    assert w2.shape == (512, hidden) and np.all(w2[384:] == 0)
    np.testing.assert_allclose(w2[:384], dense["down"], atol=2e-3, rtol=2e-3)


def test_pad_dim_matches_reference_pad_semantics():
    t = torch.arange(6).reshape(2, 3)
    assert pad_dim(t, 0, 1).shape == (2, 3) and pad_dim(t, 4, 1).shape == (2, 4) and pad_dim(t, 3, 1).shape == (2, 3)
    assert pad_dim(t, 4, 0).shape == (4, 3) and int(pad_dim(t, 4, 0)[2:].abs().sum()) == 0


def test_unsupported_bit_widths_are_rejected_loudly(tmp_path):
    rng = np.random.default_rng(1)
    path = str(tmp_path / "l.safetensors")
    names, _ = ckpt_util.write_checkpoint(path, rng, "gptq", 128, 2, 1, 64, 128)
    with pytest.raises(ValueError):
        B200Loader(CheckpointReader(path), QuantConfig("gptq", bits=3), device="cpu").groupwise_tensors([names["o"]])
