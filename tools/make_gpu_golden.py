#!/usr/bin/env python
"""Golden vectors from the reference's own CUDA ops, generated ON A GPU BOX (run under gpurun; the outputs land in
gpurun_out/ and are then committed under tests/golden/): the reference binds flashinfer's kernels for its glue ops
(rtp_llm/models_py/bindings/cuda/RegisterBaseBindings.hpp:45-160 -> 3rdparty/flashinfer/flashinfer.h:24-40):
rmsnorm, fused_add_rmsnorm, silu_and_mul. flashinfer (same major version the reference pins) is in the image; these are the
kernels the reference would run on this GPU.  Nothing in here is product code."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def main():
    import flashinfer
    dev = torch.device("cuda")
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator(device=dev).manual_seed(123)
    out = {"flashinfer_version": np.array(flashinfer.__version__)}
    for tag, dtype in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        for rows, hidden in ((4, 512), (3, 4096)):
            x = (torch.randn(rows, hidden, generator=g, device=dev) * 1.5).to(dtype)
            r = torch.randn(rows, hidden, generator=g, device=dev).to(dtype)
            w = (1 + 0.2 * torch.randn(hidden, generator=g, device=dev)).to(dtype)
            key = f"{tag}_{hidden}"
            out[f"{key}_x"], out[f"{key}_res"], out[f"{key}_w"] = [t.view(torch.int16).cpu().numpy() for t in (x, r, w)]
            y = flashinfer.norm.rmsnorm(x.clone(), w, eps=1e-6)
            out[f"{key}_rmsnorm"] = y.view(torch.int16).cpu().numpy()
            xi, ri = x.clone(), r.clone()
            flashinfer.norm.fused_add_rmsnorm(xi, ri, w, eps=1e-6)          # in place: xi = normalised, ri = x + r
            out[f"{key}_fused_y"], out[f"{key}_fused_res"] = xi.view(torch.int16).cpu().numpy(), ri.view(torch.int16).cpu().numpy()
        gu = (torch.randn(5, 2 * 768, generator=g, device=dev) * 2).to(dtype)
        out[f"{tag}_gate_up"] = gu.view(torch.int16).cpu().numpy()
        out[f"{tag}_silu_and_mul"] = flashinfer.activation.silu_and_mul(gu).view(torch.int16).cpu().numpy()
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(OUT, "flashinfer_glue_ops.npz"), **out)
    print("wrote", os.path.join(OUT, "flashinfer_glue_ops.npz"), "flashinfer", flashinfer.__version__)


if __name__ == "__main__":
    main()
