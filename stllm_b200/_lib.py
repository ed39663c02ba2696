"""ctypes binding of the C-ABI library (include/stllm_b200.h).

There is no fallback: if ``libstllm_b200.so`` is missing or an entry point fails, an exception is raised.
Build the library with ``python __graft_entry__.py`` (or ``make -C stllm_b200/csrc``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstllm_b200.so")

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
c_ll_p = ctypes.POINTER(ctypes.c_longlong)

# name -> argtypes ; every function returns int except stllm_last_error
SIGNATURES = {
    "stllm_abi_version": [],
    "stllm_device_sm_count": [],
    "stllm_gemm_workspace_bytes": [],
    "stllm_set_gemm_workspace": [c_void_p, c_ll],
    "stllm_gemm_debug_counters": [c_void_p],
    "stllm_gemm_debug_trace": [c_int, c_void_p, c_void_p],
    "stllm_set_gemm_streamk": [c_int],
    "stllm_set_attention_variant": [c_int, c_int],
    "stllm_attention_last_launches": [],
    "stllm_gemm_plan": [c_int, c_int, c_int, c_int, c_ll_p],
    "stllm_set_attention_tc": [c_int],
    "stllm_gemm_f16": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_ll,
                       c_int, c_int, c_int, c_void_p],
    "stllm_gemm_norm_f16": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p,
                            c_ll, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_float, c_void_p, c_int,
                            c_int, c_void_p],
    "stllm_attention_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll_p, c_ll_p,
                            c_ll_p, c_ll_p, c_float, c_float, c_int, c_void_p, c_float, c_void_p],
    "stllm_layernorm_f16": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p],
    "stllm_rmsnorm_f16": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_float, c_void_p],
    "stllm_im2col_patch": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "stllm_vit_assemble_f16": [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "stllm_rope_f16": [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "stllm_gather_rows_f16": [c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_int, c_void_p,
                              c_void_p],
    "stllm_gather_rows2_f16": [c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int,
                               c_void_p, c_void_p],
    "stllm_mean_frames_f16": [c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p],
    "stllm_residual_merge_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_void_p],
    "stllm_preprocess_frames_u8": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                   c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_void_p,
                                   c_void_p],
    "stllm_bt_vit_to_branch_f16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                   c_int, c_void_p],
    "stllm_bt_branch_to_vit_avg_f16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "stllm_cosine_loss_f16": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p],
    "stllm_cross_entropy_f16": [c_void_p, c_ll, c_void_p, c_int, c_int, c_ll, c_void_p, c_void_p],
    "stllm_decode_linear_f16": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p,
                                c_float, c_void_p, c_ll, c_int, c_void_p],
    "stllm_decode_rope_append_f16": [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_ll, c_ll, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "stllm_decode_attention_workspace_bytes": [c_int, c_int],
    "stllm_decode_attention_fused_workspace_bytes": [c_int, c_int],
    "stllm_decode_attention_fused_f16": [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll, c_void_p,
                                         c_int, c_void_p, c_void_p, c_ll, c_ll, c_int, c_void_p, c_int, c_void_p,
                                         c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_float, c_void_p, c_ll,
                                         c_void_p, c_ll, c_void_p],
    "stllm_decode_attention_f16": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_int, c_void_p, c_void_p, c_ll,
                                   c_ll, c_int, c_void_p, c_int, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int,
                                   c_int, c_int, c_float, c_void_p],
}


class StllmNativeError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once). Raises StllmNativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StllmNativeError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python __graft_entry__.py` or "
            f"`make -C stllm_b200/csrc`). There is no CPU/PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.stllm_gemm_workspace_bytes.restype = c_ll
    lib.stllm_decode_attention_workspace_bytes.restype = c_ll
    lib.stllm_decode_attention_fused_workspace_bytes.restype = c_ll
    lib.stllm_last_error.argtypes = []
    lib.stllm_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().stllm_last_error().decode("utf-8", "replace")
        raise StllmNativeError(f"{what} failed with code {rc}: {msg}")
