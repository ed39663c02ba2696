"""GPU: decode-step kernels (csrc/decode.cu) against torch fp32 statements of the same ops, and the decode engine
(CUDA-graph step, shared prompt KV, ancestry table) against the general-kernel cache route, through the C ABI."""
import pytest

pytestmark = pytest.mark.gpu


def test_decode_kernels_and_engine():
    import sys
    from scripts import gpu_check_decode as K
    argv, sys.argv = sys.argv, ["gpu_check_decode.py", "--no-time"]
    try:
        K.fails.clear()
        rc = K.main()
    finally:
        sys.argv = argv
    assert rc == 0, f"decode checks failed: {K.fails}"
