"""Whole-decode-step parity: the GPU step (rtp_llm_b200.decode_step.DecodeStep on the TINY config) against the CPU oracle
composed op by op on the same weights / page tables / inputs. Used by tests (-m gpu) and __graft_entry__.smoke()."""
import dataclasses

import numpy as np
import torch

from oracle import oracle as orc
from rtp_llm_b200.decode_step import TINY, DecodeStep


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def _gemm(x_bits, ref, is_bf16=False):
    fmt, w, s, zs = ref
    if fmt == "int4":
        return orc.dequant_gemm(x_bits, "int4", w.numpy(), scales=_bits(s), zeros_x_scales=_bits(zs), group=128, is_bf16=is_bf16)
    if fmt == "int8g":
        return orc.dequant_gemm(x_bits, "int8g", w.numpy(), scales=_bits(s), zeros_x_scales=_bits(zs), group=128, is_bf16=is_bf16)
    if fmt == "int8":
        return orc.dequant_gemm(x_bits, "int8", w.numpy(), scales=_bits(s), is_bf16=is_bf16)
    return orc.dequant_gemm(x_bits, "f16", _bits(w), is_bf16=is_bf16)


def oracle_step(model: DecodeStep, kv_before):
    cfg = model.cfg
    ids = model.ids_h.numpy()
    seq = model.seq_lens_h.numpy()
    pl = orc.convert_block_table(model.block_ids_h.numpy())
    resid = _bits(model.embed)[ids]
    proj = None
    for li, L in enumerate(model.layers):
        if li == 0:
            x, _ = orc.add_rmsnorm(resid, None, _bits(L["ln1"]), cfg.eps)
        else:
            x, resid = orc.add_rmsnorm(proj, resid, _bits(L["ln1"]), cfg.eps)
        qkv = _gemm(x, L["ref"]["qkv"])
        q, pool = orc.rope_append(qkv, kv_before[li], pl, seq, model.Hq, model.Hkv, model.D, cfg.tokens_per_block, cfg.rope_base)
        a = orc.paged_decode_attn(q.reshape(model.B, model.Hq, model.D), pool, pl, seq, model.Hq, model.Hkv, model.D,
                                  cfg.tokens_per_block)
        proj = _gemm(a, L["ref"]["o"])
        x, resid = orc.add_rmsnorm(proj, resid, _bits(L["ln2"]), cfg.eps)
        gu = _gemm(x, L["ref"]["w13"])
        act = orc.silu_and_mul(gu)
        proj = _gemm(act, L["ref"]["w2"])
    x, resid = orc.add_rmsnorm(proj, resid, _bits(model.final_ln), cfg.eps)
    logits = _gemm(x, model.lm_head_ref)
    return orc.from_bits(logits, False)


def check_tiny_step(dev, quant="int4", batch=3, ctx=40, graph=False, pdl=False, program=False, fuse_rope=False) -> float:
    cfg = dataclasses.replace(TINY, quant=quant)
    model = DecodeStep(cfg, batch, ctx, dev, keep_reference=True, ragged=True, seed=1, pdl=pdl, fuse_rope=fuse_rope)
    kv_before = [_bits(L["kv"]) for L in model.layers]
    if program:
        model.build_program()
        assert model.prog.num_ops > 0, "nothing was fused into the persistent kernel"
    if graph:
        for L, kb in zip(model.layers, kv_before):   # capture() runs the step (appends K/V): the append is idempotent
            pass
        model.capture()
        model.replay()
    else:
        model.run()
    torch.cuda.synchronize(dev)
    if pdl:
        from rtp_llm_b200 import ops
        ops.set_pdl(False)
    got = model.logits.float().cpu().numpy()
    exp = oracle_step(model, kv_before)
    scale = float(np.sqrt((exp ** 2).mean()))
    err = float(np.abs(got - exp).max())
    assert np.isfinite(got).all()
    assert err <= 3e-2 * scale + 3e-2, f"logits differ: max err {err} vs rms {scale}"
    # greedy token: must equal the oracle's argmax wherever the oracle's top-2 margin exceeds the error bound
    nxt = model.next_ids.cpu().numpy()
    oa = orc.argmax(exp)
    srt = np.sort(exp, axis=-1)
    margin = srt[:, -1] - srt[:, -2]
    sure = margin > 2 * err + 1e-6
    assert (nxt[sure] == oa[sure]).all(), (nxt, oa)
    return err
