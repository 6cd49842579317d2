"""-m gpu: every CUDA entry point, called through the C ABI, against the numpy oracle (tests/kernel_cases.py).
Tolerances: bf16 storage <= 1e-2 of the tensor's max magnitude (north_star), fp32 paths <= 1e-5 .. 2e-4."""
import pytest

import kernel_cases as K


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(K.CASES))
def test_kernel_case(name):
    r = K.CASES[name]()
    assert r["ok"], f"{name}: {r}"
