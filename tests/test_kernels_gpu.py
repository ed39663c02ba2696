"""GPU: every C-ABI op against a plain torch fp32 statement of the same op (scripts/gpu_check_kernels.py holds the case
list: ragged M/N/K, every epilogue, every tile engine incl. stream-K and swapped operands, folded norms with every
producer tile width, both attention kernels).  One pytest item per case, so a regression is located by name.
Tolerances are per op class and stated where they are set (north_star's rtol 1e-3 / atol 1e-4 wherever one fp16
rounding separates the kernel from the fp32 reference)."""
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    from scripts import gpu_check_kernels as K
    return K.build_cases()


def pytest_generate_tests(metafunc):
    if "kernel_case" in metafunc.fixturenames:
        import torch
        names = [n for n, _ in _cases()] if torch.cuda.is_available() else ["needs-gpu"]
        metafunc.parametrize("kernel_case", names)


def test_kernel_vs_torch_fp32(kernel_case):
    from scripts import gpu_check_kernels as K
    fn = dict(K.build_cases())[kernel_case]
    K.fails.clear()
    assert fn(), f"{kernel_case}: {K.fails}"
