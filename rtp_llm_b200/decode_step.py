"""One decode step of a dense Llama/Qwen-style stack on the B200 hot path (synthetic weights, synthetic page tables).

This is the measurement harness of SURVEY.md section 8(d): per layer
    add_rmsnorm -> qkv GEMM -> rope + KV append -> paged decode attention -> o GEMM -> [TP all-reduce]
    -> add_rmsnorm -> w13 GEMM -> SiLU*mul -> w2 GEMM -> [TP all-reduce]
then final norm -> lm_head (fp16 weights) -> greedy argmax, all on one stream under one CUDA graph.
It mirrors the call order of the reference's Python model graph (model_desc/qwen3.py:57-138,
modules/hybrid/causal_attention.py:75-93, modules/hybrid/dense_mlp.py:95-106, PyWrappedModel.cc:938-1060) but owns none
of its engine: scheduler, cache manager and sampler policy stay the reference's.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import B200_FMT_F16, B200_FMT_INT4, B200_FMT_INT8, B200_FMT_INT8G


@dataclasses.dataclass
class ModelConfig:
    name: str
    hidden: int
    layers: int
    head_num: int
    kv_head_num: int
    head_dim: int
    inter: int
    vocab: int
    rope_base: float = 500000.0
    eps: float = 1e-5
    quant: str = "int4"          # "f16" | "int8" (per column) | "int4" (group 128) | "int8g" (8-bit group 128)
    tokens_per_block: int = 64   # KVCacheConfig.seq_size_per_block default (ConfigModules.h:174)


LLAMA3_8B = ModelConfig("Llama-3-8B", 4096, 32, 32, 8, 128, 14336, 128256)
QWEN2_72B = ModelConfig("Qwen2-72B", 8192, 80, 64, 8, 128, 29696, 152064, rope_base=1000000.0, eps=1e-6)  # GPTQ-padded inter
TINY = ModelConfig("tiny", 512, 2, 4, 2, 128, 512, 1024, tokens_per_block=16)


def _fmt(q: str) -> int:
    return {"f16": B200_FMT_F16, "int8": B200_FMT_INT8, "int4": B200_FMT_INT4, "int8g": B200_FMT_INT8G}[q]


def weight_bytes(cfg: ModelConfig, K: int, N: int) -> float:
    """Algorithmic HBM bytes of one weight (int4 = E/2 + (E/128)*4, int8 = E + 2N, f16 = 2E)."""
    E = K * N
    if cfg.quant == "int4":
        return E / 2 + E / 128 * 4
    if cfg.quant == "int8":
        return E + 2 * N
    if cfg.quant == "int8g":
        return E + E / 128 * 4
    return 2 * E


class DecodeStep:
    """Synthetic decode step. tp_size > 1 shards exactly as the reference's TP split (utils/model_weight.py:1489-1580):
    qkv / w13 column-parallel, o / w2 row-parallel + all-reduce, lm_head vocab-parallel + all-gather."""

    def __init__(self, cfg: ModelConfig, batch: int, ctx: int, device: torch.device, tp_rank: int = 0, tp_size: int = 1,
                 dtype=torch.float16, seed: int = 0, keep_reference: bool = False, ragged: bool = False,
                 pdl: bool = False, comm=None, fuse_silu: bool = True, fuse_ar_norm: bool = True, fuse_rope: Optional[bool] = None,
                 fuse_gemm_rs: Optional[bool] = None):
        assert cfg.head_num % tp_size == 0 and cfg.inter % tp_size == 0
        self.cfg, self.B, self.ctx, self.dev, self.dtype = cfg, batch, ctx, device, dtype
        self.tp_rank, self.tp_size, self.comm, self.pdl = tp_rank, tp_size, comm, pdl
        self.Hq = cfg.head_num // tp_size
        self.Hkv = max(cfg.kv_head_num // tp_size, 1)   # kv heads replicated when Hkv < tp (MHAKVCacheSpec.h:46-51)
        self.D = cfg.head_dim
        self.inter = cfg.inter // tp_size
        assert (self.inter % 128 == 0) and (self.Hq * self.D) % 128 == 0, "row-parallel K must be a multiple of 128"
        self.vocab = (cfg.vocab + tp_size - 1) // tp_size
        self.vocab = (self.vocab + 7) // 8 * 8          # sp_0_pad8
        T = cfg.tokens_per_block
        self.M = (ctx + T - 1) // T
        self.ref: Dict[str, list] = {} if keep_reference else None
        self.fuse_silu = fuse_silu and self.inter % 64 == 0
        self.fuse_ar_norm = fuse_ar_norm
        # row-parallel GEMM + reduce-scatter in one kernel (b200_wo_gemm_rs), then gather + residual + norm (b200_peer_gather_norm):
        # only with the peer communicator and the fused all-reduce+norm; B200_FUSE_GEMM_RS=0 turns it off (A/B runs)
        if fuse_gemm_rs is None:
            fuse_gemm_rs = os.environ.get("B200_FUSE_GEMM_RS", "1") == "1"
        self.fuse_gemm_rs = (bool(fuse_gemm_rs) and fuse_ar_norm and tp_size > 1 and hasattr(comm, "gemm_rs")
                             and comm.gemm_rs_supported(batch, cfg.hidden))
        # RoPE + K/V append inside the attention kernel (b200_paged_decode_attn_rope): bit-identical, one launch less per layer,
        # but measured 0.4 % SLOWER at the headline config (same box A/B, profiles/r02_fuse_rope_ab.txt): under PDL the stand-alone
        # rope kernel hides behind its neighbours while the fused prologue sits on the attention kernel's critical path.
        # Default off; B200_FUSE_ROPE=1 or fuse_rope=True turns it on.
        if fuse_rope is None:
            fuse_rope = os.environ.get("B200_FUSE_ROPE", "0") == "1"
        self.fuse_rope = bool(fuse_rope) and cfg.head_dim == 128
        g = torch.Generator(device="cpu").manual_seed(seed * 1000 + 17)
        H = cfg.hidden

        def pack_ref(refw, gate_up_inter: int = 0):
            """Reference-layout tuple -> kernel layout. gate_up_inter > 0: a [K, 2I] gate|up weight whose columns are
            first interleaved per 64 outputs so the GEMM epilogue can apply SiLU(gate)*up (ops.interleave_gate_up)."""
            fmt, w, s, zs = refw
            gi = gate_up_inter
            if fmt == "int4":
                w, s, zs = w.to(device), s.to(device), zs.to(device)
                if gi:
                    w, s, zs = (ops.interleave_gate_up(w, gi, packed_int4=True), ops.interleave_gate_up(s, gi),
                                ops.interleave_gate_up(zs, gi))
                return ops.pack_w4(w, s, zs)
            if fmt == "int8g":
                w, s, zs = w.to(device), s.to(device), zs.to(device)
                if gi:
                    w, s, zs = ops.interleave_gate_up(w, gi), ops.interleave_gate_up(s, gi), ops.interleave_gate_up(zs, gi)
                return ops.pack_w8g(w, s, zs)
            if fmt == "int8":
                w, s = w.to(device), s.to(device)
                if gi:
                    w, s = ops.interleave_gate_up(w, gi), ops.interleave_gate_up(s, gi)
                return ops.pack_w8(w, s)
            w = w.to(device)
            if gi:
                w = ops.interleave_gate_up(w, gi)
            return ops.pack_f16(w)

        def make_weight(K: int, N: int, gen_seed: int, quant: Optional[str] = None, gate_up_inter: int = 0):
            quant = quant or cfg.quant
            gg = torch.Generator(device=device).manual_seed(gen_seed)
            if quant == "int4":
                qp = torch.randint(0, 256, (K, N // 2), generator=gg, device=device, dtype=torch.uint8)
                s = (torch.randn(K // 128, N, generator=gg, device=device).abs() * 0.01 + 1e-3).to(dtype)
                z = torch.randint(0, 16, (K // 128, N), generator=gg, device=device)
                zs = ((8 - z).to(dtype) * s).to(dtype)
                refw = ("int4", qp, s, zs)
            elif quant == "int8g":
                q8 = torch.randint(-128, 128, (K, N), generator=gg, device=device, dtype=torch.int8)
                s = (torch.randn(K // 128, N, generator=gg, device=device).abs() * 6e-4 + 6e-5).to(dtype)
                z = torch.randint(0, 256, (K // 128, N), generator=gg, device=device)
                zs = ((128 - z).to(dtype) * s).to(dtype)
                refw = ("int8g", q8, s, zs)
            elif quant == "int8":
                q8 = torch.randint(-128, 128, (K, N), generator=gg, device=device, dtype=torch.int8)
                s = (torch.randn(N, generator=gg, device=device).abs() * 2e-4 + 1e-5).to(dtype)
                refw = ("int8", q8, s, None)
            else:
                wkn = (torch.randn(K, N, generator=gg, device=device) * 0.02).to(dtype)
                refw = ("f16", wkn, None, None)
            w = pack_ref(refw, gate_up_inter)
            return w, (tuple(t.cpu() if torch.is_tensor(t) else t for t in refw) if keep_reference else None)

        self.layers: List[dict] = []
        qkv_n = (self.Hq + 2 * self.Hkv) * self.D
        from . import tp as tpmod
        shard_full = keep_reference and tp_size > 1      # parity runs: every rank builds the FULL weights, then slices
        for l in range(cfg.layers):
            base = (seed * 100003 + l) * 16 + (0 if shard_full else tp_rank * 7919)
            L = {}
            if shard_full:
                _, f0 = make_weight(H, (cfg.head_num + 2 * cfg.kv_head_num) * self.D, base + 1)
                _, f1 = make_weight(cfg.head_num * self.D, H, base + 2)
                _, f2 = make_weight(H, 2 * cfg.inter, base + 3)
                _, f3 = make_weight(cfg.inter, H, base + 4)
                r0 = tpmod.shard_qkv(f0, cfg.head_num, cfg.kv_head_num, self.D, tp_rank, tp_size)
                r1 = tpmod.shard_o(f1, cfg.head_num, self.D, tp_rank, tp_size)
                r2 = tpmod.shard_w13(f2, cfg.inter, tp_rank, tp_size)
                r3 = tpmod.shard_w2(f3, cfg.inter, tp_rank, tp_size)
                L["qkv"], L["o"], L["w2"] = pack_ref(r0), pack_ref(r1), pack_ref(r3)
                L["w13"] = pack_ref(r2, self.inter if self.fuse_silu else 0)
                L["full"] = dict(qkv=f0, o=f1, w13=f2, w2=f3)
            else:
                L["qkv"], r0 = make_weight(H, qkv_n, base + 1)
                L["o"], r1 = make_weight(self.Hq * self.D, H, base + 2)
                L["w13"], r2 = make_weight(H, 2 * self.inter, base + 3, gate_up_inter=self.inter if self.fuse_silu else 0)
                L["w2"], r3 = make_weight(self.inter, H, base + 4)
            L["ln1"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(dtype).to(device)
            L["ln2"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(dtype).to(device)
            L["kv"] = None
            if keep_reference:
                L["ref"] = dict(qkv=r0, o=r1, w13=r2, w2=r3)
            self.layers.append(L)
        self.final_ln = (1.0 + 0.1 * torch.randn(H, generator=g)).to(dtype).to(device)
        if shard_full:
            _, self.lm_head_full = make_weight(H, cfg.vocab, seed * 100003 + 999983, quant="f16")
            self.lm_head_ref = tpmod.shard_lm_head(self.lm_head_full, self.vocab, tp_rank)
            self.lm_head = pack_ref(self.lm_head_ref)
        else:
            self.lm_head, self.lm_head_ref = make_weight(H, self.vocab, seed * 100003 + 999983 + tp_rank, quant="f16")
        gg = torch.Generator(device=device).manual_seed(seed + 5)
        self.embed = (torch.randn(cfg.vocab, H, generator=gg, device=device) * 0.5).to(dtype)

        # ---- paged KV cache: P = B*M + 1 pages per layer, block 0 reserved, random permutation of the rest
        P = batch * self.M + 1
        gk = torch.Generator(device=device).manual_seed(42)
        for L in self.layers:
            if shard_full:
                # parity runs: the FULL pool (all kv heads) from a common seed, then this rank's heads -- every rank holds
                # DIFFERENT K/V, so a kv-head mix-up between ranks cannot cancel out against the unsharded oracle
                full = torch.randn(P, 2, cfg.kv_head_num, T, self.D, generator=gk, device=device).to(dtype)
                kv_rank = tp_rank if cfg.kv_head_num >= tp_size else tp_rank * cfg.kv_head_num // tp_size
                L["kv_full"] = full.cpu()
                L["kv"] = full[:, :, kv_rank * self.Hkv:(kv_rank + 1) * self.Hkv].contiguous()
            else:
                L["kv"] = torch.randn(P, 2, self.Hkv, T, self.D, generator=gk, device=device).to(dtype)
        gp = torch.Generator().manual_seed(3)
        perm = torch.randperm(P - 1, generator=gp).to(torch.int32) + 1
        self.block_ids_h = perm.reshape(batch, self.M).contiguous().pin_memory()
        if ragged:
            gl = torch.Generator().manual_seed(4)
            lens = torch.randint(ctx // 2, ctx + 1, (batch,), generator=gl)
        else:
            lens = torch.full((batch,), ctx)
        self.seq_lens_h = (lens - 1).to(torch.int32).pin_memory()       # tokens already cached
        self.ids_h = torch.randint(0, cfg.vocab, (batch,), generator=gp).to(torch.int32).pin_memory()
        self.next_h = torch.zeros(batch, dtype=torch.int32).pin_memory()

        # ---- static device buffers (graph-capturable: nothing is allocated inside step())
        def buf(*shape, dt=dtype):
            return torch.empty(*shape, dtype=dt, device=device)
        self.ids = buf(batch, dt=torch.int32)
        self.seq_lens = buf(batch, dt=torch.int32)
        self.block_ids = buf(batch, self.M, dt=torch.int32)
        self.page_list = buf(batch, 1, 2, self.M, dt=torch.int32)
        self.resid = buf(batch, H)
        self.x = buf(batch, H)
        self.qkv = buf(batch, qkv_n)
        self.q = buf(batch, self.Hq * self.D)
        self.attn = buf(batch, self.Hq * self.D)
        self.proj = buf(batch, H)
        self.gu = buf(batch, 2 * self.inter)
        self.act = buf(batch, self.inter)
        self.logits = buf(batch, self.vocab)
        self.logits_all = buf(tp_size, batch, self.vocab) if tp_size > 1 else None
        self.next_ids = buf(batch, dt=torch.int32)
        shapes = [(H, qkv_n), (self.Hq * self.D, H), (H, 2 * self.inter), (self.inter, H), (H, self.vocab)]
        self.gemm_ws = ops.gemm_workspace(batch, shapes, device)
        self.attn_ws = ops.attn_workspace(batch, self.Hq, self.Hkv, ctx, device)
        ops.set_pdl(pdl)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.prog = None
        self.upload_inputs()
        torch.cuda.synchronize(device)

    # ------------------------------------------------------------------ host <-> device
    def upload_inputs(self):
        """H2D of one step's inputs from pinned host memory (token ids, lengths, block table)."""
        self.ids.copy_(self.ids_h, non_blocking=True)
        self.seq_lens.copy_(self.seq_lens_h, non_blocking=True)
        self.block_ids.copy_(self.block_ids_h, non_blocking=True)

    def download_outputs(self):
        self.next_h.copy_(self.next_ids, non_blocking=True)

    def h2d_bytes(self) -> int:
        return self.ids_h.numel() * 4 + self.seq_lens_h.numel() * 4 + self.block_ids_h.numel() * 4

    def d2h_bytes(self) -> int:
        return self.next_h.numel() * 4

    # ------------------------------------------------------------------ the step
    def _all_reduce(self, t: torch.Tensor):
        if self.tp_size > 1:
            self.comm.all_reduce(t)

    def _row_gemm(self, x, w):
        """self.proj = x . w for a row-parallel weight (o, w2). Fused form (b200_wo_gemm_rs): the GEMM epilogue pushes the
        reduce-scatter words over NVLink; the _ar_norm() that follows then only reduces / gathers / normalises."""
        if self.fuse_gemm_rs and hasattr(self.comm, "gemm_rs"):    # (bench.py swaps the communicator for its NCCL parity arm)
            self.comm.gemm_rs(x, w, self.gemm_ws, self.proj, pdl=self.pdl)
        else:
            ops.wo_gemm(x, w, self.gemm_ws, out=self.proj, pdl=self.pdl)

    def _ar_norm(self, t, gamma):
        """[TP all-reduce of the row-parallel GEMM output t] + residual add + RMSNorm -> self.x. With the peer communicator the
        three steps are ONE kernel (b200_peer_allreduce_norm); otherwise all-reduce and fused_add_rmsnorm run separately."""
        if self.fuse_gemm_rs and hasattr(self.comm, "gemm_rs"):      # t came out of _row_gemm(): its scatter phase is already on the wire
            self.comm.gather_norm(t, self.resid, gamma, self.cfg.eps, self.x)
            return
        if self.tp_size > 1:
            fused = getattr(self.comm, "all_reduce_norm", None)
            if self.fuse_ar_norm and fused is not None and fused(t, self.resid, gamma, self.cfg.eps, self.x):
                return
            self.comm.all_reduce(t)
        ops.add_rmsnorm(t, self.resid, gamma, self.cfg.eps, out=self.x)

    def step_core(self):
        """Everything up to the sampled token (TP1 / peer communicator) or the local logits (NCCL): C-ABI calls only, so it
        can be recorded into a decode program."""
        cfg = self.cfg
        ops.convert_block_table(self.block_ids, out=self.page_list)
        ops.embedding(self.ids, self.embed, out=self.resid)
        first = True
        for L in self.layers:
            if first:
                ops.add_rmsnorm(self.resid, None, L["ln1"], cfg.eps, out=self.x)
                first = False
            else:
                self._ar_norm(self.proj, L["ln1"])          # all-reduce of the previous layer's w2 output rides along
            ops.wo_gemm(self.x, L["qkv"], self.gemm_ws, out=self.qkv, pdl=self.pdl)
            if self.fuse_rope:    # RoPE + K/V append inside the attention kernel (one launch instead of two)
                ops.paged_decode_attn_rope(self.qkv, L["kv"], self.page_list, self.seq_lens, self.Hq, self.ctx, cfg.rope_base,
                                           self.attn_ws, out=self.attn)
            else:
                ops.rope_append(self.qkv, L["kv"], self.page_list, self.seq_lens, self.Hq, cfg.rope_base, q_out=self.q)
                ops.paged_decode_attn(self.q, L["kv"], self.page_list, self.seq_lens, self.ctx, self.attn_ws, out=self.attn)
            self._row_gemm(self.attn, L["o"])
            self._ar_norm(self.proj, L["ln2"])
            if self.fuse_silu:     # SiLU(gate)*up in the GEMM epilogue (gate/up columns interleaved at load time)
                ops.wo_gemm(self.x, L["w13"], self.gemm_ws, out=self.act, pdl=self.pdl, silu_mul=True)
            else:
                ops.wo_gemm(self.x, L["w13"], self.gemm_ws, out=self.gu, pdl=self.pdl)
                ops.silu_and_mul(self.gu, out=self.act)
            self._row_gemm(self.act, L["w2"])
        self._ar_norm(self.proj, self.final_ln)
        ops.wo_gemm(self.x, self.lm_head, self.gemm_ws, out=self.logits, pdl=self.pdl)
        if self.tp_size == 1:
            ops.argmax(self.logits, out=self.next_ids)
        elif hasattr(self.comm, "argmax"):
            self.comm.argmax(self.logits, cfg.vocab, self.next_ids)     # vocab-parallel greedy sampling, own kernel

    def step_tail(self):
        """TP > 1 with stock NCCL: gather the vocab-split logits and sample (torch plumbing, outside the program)."""
        if self.tp_size > 1 and not hasattr(self.comm, "argmax"):
            self.comm.all_gather(self.logits_all, self.logits)
            full = self.logits_all.permute(1, 0, 2).reshape(self.B, -1)[:, : self.cfg.vocab]   # drop the sp_0_pad8 columns
            self.next_ids.copy_(torch.argmax(full.float(), dim=-1).to(torch.int32))

    def step(self):
        """One decode step, op by op (one kernel per call)."""
        self.step_core()
        self.step_tail()

    def build_program(self):
        """Record step_core() into a decode program (b200_program_*): the ops between two attention calls become one
        persistent-kernel launch. Requires every call in step_core to go through the C ABI (own peer all-reduce, not NCCL)."""
        from . import tp as tpmod
        assert self.tp_size == 1 or isinstance(self.comm, tpmod.PeerComm), "a program cannot record NCCL calls"
        self.prog = ops.Program()
        with self.prog.record():
            self.step_core()
        return self.prog

    def run(self):
        """One decode step through the program if one was built, else op by op."""
        if getattr(self, "prog", None) is not None:
            self.prog.launch()
            self.step_tail()
        else:
            self.step()

    def launches_per_step(self) -> int:
        n0 = ops.launch_count()
        self.run()
        return ops.launch_count() - n0

    def capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self.run()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.run()
        torch.cuda.synchronize(self.dev)

    def replay(self):
        self.graph.replay()

    # ------------------------------------------------------------------ accounting
    def algorithmic_bytes(self) -> Dict[str, float]:
        cfg, H = self.cfg, self.cfg.hidden
        qkv_n = (self.Hq + 2 * self.Hkv) * self.D
        w = (weight_bytes(cfg, H, qkv_n) + weight_bytes(cfg, self.Hq * self.D, H) + weight_bytes(cfg, H, 2 * self.inter)
             + weight_bytes(cfg, self.inter, H)) * cfg.layers
        lens = (self.seq_lens_h.to(torch.int64) + 1).sum().item()
        kv = 2.0 * lens * self.Hkv * self.D * 2 * cfg.layers
        lm = 2.0 * H * self.vocab
        return dict(weights=w, kv=kv, lm_head=lm, total=w + kv + lm)
