"""GPU: every C-ABI op against a plain torch fp32 reference of the same op (per-kernel numerics; tolerances of a
few fp16 ulps are stated per case in scripts/gpu_check_kernels.py)."""
import pytest

pytestmark = pytest.mark.gpu


def test_all_kernels_against_torch_fp32():
    import sys
    from scripts import gpu_check_kernels as K
    argv, sys.argv = sys.argv, ["gpu_check_kernels.py", "--no-time"]
    try:
        K.fails.clear()
        rc = K.main()
    finally:
        sys.argv = argv
    assert rc == 0, f"kernel checks failed: {K.fails}"
