"""GPU parity tests (pytest -m gpu): each hand-written kernel, through the C-ABI, vs a PyTorch fp32 restatement."""
import pytest
import torch

from tests.kernel_checks import CHECKS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CHECKS))
def test_kernel(name):
    r = CHECKS[name]()
    torch.cuda.synchronize()
    assert r["ok"], f"{name}: max abs err {r['err']:.4g} > tol {r['tol']:.4g} (max |ref| {r['ref']:.4g})"
