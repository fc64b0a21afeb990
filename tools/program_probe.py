#!/usr/bin/env python
"""Developer probe for the decode program (persistent kernel): runs a 1- or 2-layer tiny step op by op and as a program under
different B200_PROGRAM_FUSE_MASK settings and reports, buffer by buffer, where the two first disagree."""
import dataclasses
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtp_llm_b200.decode_step import TINY, DecodeStep  # noqa: E402

dev = torch.device("cuda:0")
BUFS = ["page_list", "resid", "x", "qkv", "q", "attn", "proj", "act", "logits"]
OPS = {1: "gemm", 2: "norm", 3: "rope", 4: "embed", 6: "blocktable"}


def run(layers, quant, batch, mask, hidden=512):
    cfg = dataclasses.replace(TINY, quant=quant, layers=layers)
    m = DecodeStep(cfg, batch, 60, dev, ragged=True, seed=2)
    m.step()
    torch.cuda.synchronize()
    ref = {b: getattr(m, b).clone() for b in BUFS}
    kv_ref = [L["kv"].clone() for L in m.layers]
    os.environ["B200_PROGRAM_FUSE_MASK"] = str(mask)
    try:
        m.build_program()
    finally:
        os.environ.pop("B200_PROGRAM_FUSE_MASK", None)
    for b in BUFS:
        if b != "page_list":
            getattr(m, b).zero_()
    m.run()
    torch.cuda.synchronize()
    out = []
    for b in BUFS:
        a, r = getattr(m, b).float(), ref[b].float()
        err = (a - r).abs().max().item()
        out.append(f"{b}={err:.3g}")
    kerr = max((L["kv"].float() - k.float()).abs().max().item() for L, k in zip(m.layers, kv_ref))
    names = "+".join(OPS[t] for t in OPS if (mask >> t) & 1) or "none"
    print(f"layers={layers} {quant} B{batch} fuse[{names}] ops={m.prog.num_ops} launches={m.prog.num_launches}: "
          + " ".join(out) + f" kv={kerr:.3g}", flush=True)


if __name__ == "__main__":
    for mask in (0, 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 6, (1 << 2) | (1 << 3) | (1 << 4) | (1 << 6), (1 << 1) | (1 << 2), -1 & 0x7f):
        run(1, "int4", 3, mask)
    run(2, "int4", 3, 0x7f)
    run(2, "int8", 17, 0x7f)
    run(2, "int4", 40, 0x7f)
