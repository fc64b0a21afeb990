import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.step_oracle import check_tiny_step  # noqa: E402


@pytest.mark.parametrize("quant", ["int4", "int8", "f16"])
def test_tiny_decode_step_matches_oracle(quant):
    check_tiny_step(torch.device("cuda:0"), quant=quant, batch=3, ctx=40)


def test_tiny_decode_step_under_cuda_graph():
    check_tiny_step(torch.device("cuda:0"), quant="int4", batch=5, ctx=70, graph=True)


def test_tiny_decode_step_with_programmatic_dependent_launch():
    check_tiny_step(torch.device("cuda:0"), quant="int4", batch=5, ctx=70, graph=True, pdl=True)
    check_tiny_step(torch.device("cuda:0"), quant="int8", batch=3, ctx=40, graph=False, pdl=True)
