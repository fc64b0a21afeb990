"""Multi-GPU tests (run when the box shows >= 2 GPUs, skipped otherwise): spawn one process per GPU with torchrun and check
the peer-memory collectives against torch references and the tensor-parallel tiny decode step against the UNSHARDED CPU
oracle (every rank holds different K/V: the full pool is generated from a common seed and sliced by kv head)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc, env=None, timeout=420):
    e = dict(os.environ)
    e.update(env or {})
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    r = subprocess.run(cmd, env=e, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    lines = [l for l in r.stdout.splitlines() if l.startswith("[PASS]") or l.startswith("[FAIL]")]
    assert r.returncode == 0 and lines and not any(l.startswith("[FAIL]") for l in lines), r.stdout[-3000:] + r.stderr[-3000:]
    return lines


def _world():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    return 2


def test_peer_collectives_match_torch():
    _torchrun("tp_collectives_check.py", _world())


@pytest.mark.parametrize("gemm_rs", ["1", "0"], ids=["gemm_rs", "gemm_then_allreduce"])
@pytest.mark.parametrize("program", ["0", "1"], ids=["op_by_op", "program"])
def test_tp_tiny_step_matches_unsharded_oracle(program, gemm_rs):
    """gemm_rs=1 (default): row-parallel GEMMs push their reduce-scatter words from the epilogue (b200_wo_gemm_rs) and
    b200_peer_gather_norm finishes the exchange; gemm_rs=0: b200_wo_gemm followed by b200_peer_allreduce_norm."""
    _torchrun("tp_check.py", _world(), env={"TP_PROGRAM": program, "B200_FUSE_GEMM_RS": gemm_rs})
