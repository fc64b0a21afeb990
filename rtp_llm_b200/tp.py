"""Tensor-parallel collectives of the decode path: SUM all-reduce of [B, hidden] after each row-parallel GEMM
(reference: distributed/collective_torch.py:694-722, call sites hybrid/causal_attention.py:91-92, dense_mlp.py:104-105)
and the all-gather of vocab-split logits (cpp/models/PyWrappedModel.cc:915-936).  One process per GPU,
torch.distributed (NCCL over NVLink 5 / NVSwitch) as plumbing; both calls are CUDA-graph capturable."""
from __future__ import annotations

import torch
import torch.distributed as dist


class NcclComm:
    """Stock NCCL collectives on the TP group (the baseline arm; see PeerComm in this module for the custom kernel)."""

    def __init__(self, device: torch.device, group=None):
        self.device, self.group = device, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_reduce(self, t: torch.Tensor) -> None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def all_gather(self, out: torch.Tensor, t: torch.Tensor) -> None:
        dist.all_gather_into_tensor(out, t, group=self.group)


def make_comm(device: torch.device, group=None):
    return NcclComm(device, group)
