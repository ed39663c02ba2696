"""stllm_b200 - B200-native (sm_100a) implementation of ST-LLM's video-token hot path.

Python host code mirrors the reference's model API (stllm/models/*.py); compute runs in hand-written CUDA
kernels behind the C ABI of include/stllm_b200.h (stllm_b200/libstllm_b200.so).
"""
__version__ = "0.1.0"
