"""Thin torch-tensor wrappers over the C ABI (include/stllm_b200.h).

torch is used only for device memory and the current CUDA stream; all arithmetic happens in the sm_100a
kernels of ``libstllm_b200.so``.  Every wrapper raises on non-CUDA tensors: there is no CPU fallback.
"""
import collections
import ctypes
import math

import torch

from . import _lib

EPI_NONE, EPI_GELU, EPI_SWIGLU, EPI_RELU = 0, 1, 2, 3

# number of kernels launched through this module (bench.py reports it as gpu_launches)
launch_count = 0
# when set to a list, every gemm() call is recorded (bench.py replays the GEMMs alone to time the dominant kernel)
gemm_recorder = None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_gemm_ws = None


def _ensure_gemm_workspace(device):
    """register the (zeroed) stream-K workspace once; torch owns the memory"""
    global _gemm_ws
    if _gemm_ws is None or _gemm_ws.device != device:
        lib = _lib.load()
        n = int(lib.stllm_gemm_workspace_bytes())
        _gemm_ws = torch.zeros(n, dtype=torch.uint8, device=device)
        _lib.check(lib.stllm_set_gemm_workspace(ctypes.c_void_p(_gemm_ws.data_ptr()), n), "stllm_set_gemm_workspace")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.StllmNativeError("stllm_b200 ops need CUDA tensors (no CPU fallback exists for the hot path)")


def _rows2d(t):
    """view a [..., C] tensor with contiguous last dim and uniform row stride as (rows, ld)"""
    assert t.stride(-1) == 1, "last dim must be contiguous"
    if t.dim() == 1:
        return 1, t.shape[0]
    t2 = t if t.dim() == 2 else t.reshape(-1, t.shape[-1]) if t.is_contiguous() else None
    if t2 is None:
        raise ValueError("tensor must be 2-D or contiguous")
    return t2.shape[0], t2.stride(0)


NORM_NONE, NORM_RMS, NORM_LN = 0, 1, 2


def gemm(a, w, bias=None, residual=None, epilogue=EPI_NONE, out=None, out_dtype=torch.float16, tile_n=0, k_real=None,
         norm=NORM_NONE, rs_in=None, col_sb=None, norm_eps=0.0, rs_out=None, rs_out_mode=0):
    """out = epi(a @ w.T + bias) (+ residual).  a: [..., K] fp16, w: [N, K] fp16 (nn.Linear layout).
    k_real: un-padded K, only used for algorithmic-FLOP accounting (patch-embed 588 of 640).

    Folded normalisation (no standalone LayerNorm / RMSNorm launch): with norm=NORM_RMS / NORM_LN, `a` holds the
    UN-normalised rows, `w` the gamma-scaled weights and rs_in the per-row statistics buffer (see row_stats_buffer)
    that the producing GEMM filled through rs_out / rs_out_mode (1: sum c^2, 2: (sum c, sum c^2)).  NORM_LN
    additionally takes col_sb [N, 2] = (sum_k w[n, k], folded bias)."""
    global launch_count
    _need_cuda(a, w, bias, residual, out, rs_in, col_sb, rs_out)
    assert a.dtype == torch.float16 and w.dtype == torch.float16
    _ensure_gemm_workspace(a.device)
    K = a.shape[-1]
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    M, lda = _rows2d(a)
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        out = torch.empty(a.shape[:-1] + (n_out,), dtype=out_dtype, device=a.device)
    Mo, ldc = _rows2d(out)
    assert Mo == M and out.shape[-1] == n_out
    ldr = 0
    if residual is not None:
        assert residual.dtype == torch.float16 and residual.shape[-1] == n_out
        Mr, ldr = _rows2d(residual)
        assert Mr == M
    if bias is not None:
        assert bias.dtype == torch.float16 and bias.numel() == N and bias.is_contiguous()
    lib = _lib.load()
    if norm == NORM_NONE and rs_out is None:
        rc = lib.stllm_gemm_f16(_ptr(a), lda, _ptr(w), w.stride(0), _ptr(out), ldc, M, N, K, _ptr(bias),
                                _ptr(residual), ldr, epilogue, 1 if out.dtype == torch.float32 else 0, tile_n,
                                _stream())
        _lib.check(rc, "stllm_gemm_f16")
    else:
        if norm != NORM_NONE:
            assert rs_in.dtype == torch.float32 and rs_in.is_contiguous() and rs_in.dim() == 2 and rs_in.shape[0] == M
        if norm == NORM_LN:
            assert bias is None and col_sb.dtype == torch.float32 and col_sb.is_contiguous() and col_sb.shape == (N, 2)
        if rs_out is not None:
            assert rs_out.dtype == torch.float32 and rs_out.is_contiguous() and rs_out.dim() == 2 and rs_out.shape[0] == M
        rc = lib.stllm_gemm_norm_f16(_ptr(a), lda, _ptr(w), w.stride(0), _ptr(out), ldc, M, N, K, _ptr(bias),
                                     _ptr(residual), ldr, epilogue, 1 if out.dtype == torch.float32 else 0, tile_n,
                                     norm, _ptr(rs_in), rs_in.shape[1] if rs_in is not None else 0, _ptr(col_sb),
                                     float(norm_eps), _ptr(rs_out), rs_out_mode if rs_out is not None else 0,
                                     rs_out.shape[1] if rs_out is not None else 0, _stream())
        _lib.check(rc, "stllm_gemm_norm_f16")
    launch_count += 1
    if gemm_recorder is not None:
        gemm_recorder.append(dict(a=a, w=w, bias=bias, residual=residual, epilogue=epilogue, out=out, tile_n=tile_n,
                                  M=M, N=N, K=k_real or K, norm=norm, rs_in=rs_in, col_sb=col_sb, norm_eps=norm_eps))
    return out


def row_stats_buffer(n_buffers, rows, cols, mode, device):
    """zeroed per-row statistics buffers [n_buffers, rows, stride] for the folded norms: `cols` = width of the
    normalised rows.  The producing GEMM writes one slot of `mode` floats per (column tile, warp half) -- 2 slots per tile
    of >= 128 columns, so 2 * ceil(cols / 128) slots cover every tile width the planner may pick; the consumer sums all
    slots in order.  Row stride padded to 8 floats.  Live slots are rewritten on each forward, the others stay zero: no
    per-step zeroing is needed."""
    stride = (2 * ((cols + 127) // 128) * mode + 7) // 8 * 8
    return torch.zeros(n_buffers, rows, stride, dtype=torch.float32, device=device)


# Concurrent forwards (two pipelines replaying their CUDA graphs on two streams, bench.py --streams) must not share
# statistics buffers: the pipeline sets `stats_context` to its own id while it records / runs a step.
stats_context = 0
_stats_cache = collections.OrderedDict()
_STATS_CACHE_EAGER_MAX = 16


def cached_row_stats(owner, n_buffers, rows, cols, mode, device):
    """row_stats_buffer cached per (module, stats_context, shape): allocated once, rewritten on every forward.
    Eager calls (context 0: varying prompt lengths, decode steps) share a small LRU; buffers of a pipeline context are
    kept for the life of the process because its CUDA graph holds their addresses."""
    key = (id(owner), stats_context, n_buffers, rows, cols, mode, str(device))
    buf = _stats_cache.get(key)
    if buf is None:
        buf = _stats_cache[key] = row_stats_buffer(n_buffers, rows, cols, mode, device)
        eager = [k for k in _stats_cache if k[1] == 0]
        if len(eager) > _STATS_CACHE_EAGER_MAX:
            del _stats_cache[eager[0]]
    else:
        _stats_cache.move_to_end(key)
    return buf


def gemm_plan(M, N, K, epilogue=EPI_NONE):
    """(engine, tile columns, estimated cycles) the host cost model picks: engine 0 one-CTA, 1 CTA pair, 2 swapped"""
    out = (ctypes.c_longlong * 3)()
    _lib.check(_lib.load().stllm_gemm_plan(M, N, K, epilogue, out), "stllm_gemm_plan")
    return int(out[0]), int(out[1]), int(out[2])


def _strides3(t, layout):
    """(batch, seq, head) element strides of a 4-D tensor given as 'bshd' or 'bhsd'"""
    assert t.dim() == 4 and t.stride(-1) == 1
    if layout == "bshd":
        s = (t.stride(0), t.stride(1), t.stride(2))
    else:
        s = (t.stride(0), t.stride(2), t.stride(1))
    return (ctypes.c_longlong * 3)(*s)


def attention(q, k, v, scale, causal=False, key_mask=None, mask_value=-10000.0, q_prescale=0.0, out=None):
    """q: [B, Sq, H, D], k/v: [B, Skv, H, D] (any strides with contiguous D) -> out [B, Sq, H, D] fp16."""
    global launch_count
    _need_cuda(q, k, v, key_mask, out)
    B, Sq, H, D = q.shape
    Skv = k.shape[1]
    assert k.shape == (B, Skv, H, D) and v.shape == (B, Skv, H, D)
    if out is None:
        out = torch.empty((B, Sq, H, D), dtype=torch.float16, device=q.device)
    if key_mask is not None:
        key_mask = key_mask.to(torch.uint8).contiguous()
        assert key_mask.shape == (B, Skv)
    rc = _lib.load().stllm_attention_f16(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, H, Sq, Skv, D, _strides3(q, "bshd"),
                                         _strides3(k, "bshd"), _strides3(v, "bshd"), _strides3(out, "bshd"),
                                         float(scale), float(q_prescale), 1 if causal else 0, _ptr(key_mask),
                                         float(mask_value), _stream())
    _lib.check(rc, "stllm_attention_f16")
    launch_count += _lib.load().stllm_attention_last_launches()
    return out


def layernorm(x, gamma, beta, eps, out=None):
    global launch_count
    _need_cuda(x, gamma, beta, out)
    assert x.dtype == torch.float16 and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    rows, ldx = _rows2d(x)
    if out is None:
        out = torch.empty_like(x, memory_format=torch.contiguous_format)
    _, ldy = _rows2d(out)
    rc = _lib.load().stllm_layernorm_f16(_ptr(x), ldx, _ptr(out), ldy, _ptr(gamma), _ptr(beta), rows, x.shape[-1],
                                         float(eps), _stream())
    _lib.check(rc, "stllm_layernorm_f16")
    launch_count += 1
    return out


def rmsnorm(x, weight, eps, out=None):
    global launch_count
    _need_cuda(x, weight, out)
    assert x.dtype == torch.float16 and weight.dtype == torch.float16
    rows, ldx = _rows2d(x)
    if out is None:
        out = torch.empty_like(x, memory_format=torch.contiguous_format)
    _, ldy = _rows2d(out)
    rc = _lib.load().stllm_rmsnorm_f16(_ptr(x), ldx, _ptr(out), ldy, _ptr(weight), rows, x.shape[-1], float(eps),
                                       _stream())
    _lib.check(rc, "stllm_rmsnorm_f16")
    launch_count += 1
    return out


def im2col_patch(img, patch, kpad):
    """img [n, C, H, W] fp16/fp32 -> [n * (H/P) * (W/P), kpad] fp16"""
    global launch_count
    _need_cuda(img)
    assert img.dim() == 4 and img.is_contiguous() and img.dtype in (torch.float16, torch.float32)
    n, C, H, W = img.shape
    out = torch.empty((n * (H // patch) * (W // patch), kpad), dtype=torch.float16, device=img.device)
    rc = _lib.load().stllm_im2col_patch(_ptr(img), 1 if img.dtype == torch.float32 else 0, _ptr(out), n, C, H, W,
                                        patch, kpad, _stream())
    _lib.check(rc, "stllm_im2col_patch")
    launch_count += 1
    return out


def vit_assemble(patches, cls, pos, n_img, n_patches):
    global launch_count
    _need_cuda(patches, cls, pos)
    dim = patches.shape[-1]
    x = torch.empty((n_img, n_patches + 1, dim), dtype=torch.float16, device=patches.device)
    rc = _lib.load().stllm_vit_assemble_f16(_ptr(patches), patches.stride(0), _ptr(cls), _ptr(pos), _ptr(x), n_img,
                                            n_patches, dim, _stream())
    _lib.check(rc, "stllm_vit_assemble_f16")
    launch_count += 1
    return x


def rope_(x2d, pos, cos_t, sin_t, nheads, head_dim):
    """in-place RoPE on the first nheads*head_dim columns of x2d [rows, ld]"""
    global launch_count
    _need_cuda(x2d, pos, cos_t, sin_t)
    assert x2d.dim() == 2 and x2d.stride(1) == 1 and pos.dtype == torch.int32
    assert cos_t.dtype == torch.float32 and cos_t.shape[1] == head_dim // 2 and cos_t.is_contiguous()
    rc = _lib.load().stllm_rope_f16(_ptr(x2d), x2d.stride(0), _ptr(pos), _ptr(cos_t), _ptr(sin_t), x2d.shape[0],
                                    nheads, head_dim, cos_t.shape[0], _stream())
    _lib.check(rc, "stllm_rope_f16")
    launch_count += 1
    return x2d


def gather_rows(src, idx, out=None, addend=None):
    """out[i, :] = src[idx[i], :] (+ addend[i % addend.shape[0]])   src [R, C] fp16, idx int64 [n]"""
    global launch_count
    _need_cuda(src, idx, out)
    assert src.dim() == 2 and src.stride(1) == 1 and src.dtype == torch.float16
    idx = idx.to(torch.int64).contiguous()
    n = idx.numel()
    if out is None:
        out = torch.empty((n, src.shape[1]), dtype=torch.float16, device=src.device)
    if n == 0:
        return out
    if addend is not None:
        assert addend.dtype == torch.float16 and addend.is_contiguous() and addend.shape[1] == src.shape[1]
    rc = _lib.load().stllm_gather_rows_f16(_ptr(src), src.stride(0), src.shape[0], _ptr(idx), _ptr(out),
                                           out.stride(0), n, src.shape[1], _ptr(addend),
                                           addend.shape[0] if addend is not None else 0, ctypes.c_void_p(0),
                                           _stream())
    _lib.check(rc, "stllm_gather_rows_f16")
    launch_count += 1
    return out


def gather_rows2(src0, src1, idx, out=None):
    """out[i, :] = src0[idx[i], :] if idx[i] >= 0 else src1[-1 - idx[i], :]   (fp16 [R, C] sources, int64 idx [n])"""
    global launch_count
    _need_cuda(src0, src1, idx, out)
    for s in (src0, src1):
        assert s.dim() == 2 and s.stride(1) == 1 and s.dtype == torch.float16
    assert src0.shape[1] == src1.shape[1]
    idx = idx.to(torch.int64).contiguous()
    n = idx.numel()
    if out is None:
        out = torch.empty((n, src0.shape[1]), dtype=torch.float16, device=src0.device)
    if n == 0:
        return out
    rc = _lib.load().stllm_gather_rows2_f16(_ptr(src0), src0.stride(0), src0.shape[0], _ptr(src1), src1.stride(0),
                                            src1.shape[0], _ptr(idx), _ptr(out), out.stride(0), n, src0.shape[1],
                                            ctypes.c_void_p(0), _stream())
    _lib.check(rc, "stllm_gather_rows2_f16")
    launch_count += 1
    return out


def mean_frames(x):
    """x [B, T, ...] fp16 -> [B, ...] mean over T"""
    global launch_count
    _need_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.float16
    B, T = x.shape[:2]
    inner = x[0, 0].numel()
    y = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.float16, device=x.device)
    rc = _lib.load().stllm_mean_frames_f16(_ptr(x), _ptr(y), B, T, inner, _stream())
    _lib.check(rc, "stllm_mean_frames_f16")
    launch_count += 1
    return y


def residual_merge(x, g, idx):
    """x [B, T, L, D], g [B, L, D], idx int64 [R] -> [B, R, L, D] = x[:, idx] + g[:, None]"""
    global launch_count
    _need_cuda(x, g, idx)
    assert x.is_contiguous() and g.is_contiguous()
    B, T = x.shape[:2]
    inner = x[0, 0].numel()
    idx = idx.to(torch.int64).contiguous()
    R = idx.numel()
    out = torch.empty((B, R) + tuple(x.shape[2:]), dtype=torch.float16, device=x.device)
    rc = _lib.load().stllm_residual_merge_f16(_ptr(x), _ptr(g), _ptr(idx), _ptr(out), B, T, R, inner, _stream())
    _lib.check(rc, "stllm_residual_merge_f16")
    launch_count += 1
    return out


def bt_vit_to_branch(x, B, T, P, z_prev=None, out=None, pos=None, time_emb=None, bt_cls=None):
    """ViT stream x [(B*T)*(P+1), D] -> BT-Adapter branch layout [B + B*P*T, D] (+ z_prev), see include/stllm_b200.h"""
    global launch_count
    _need_cuda(x, z_prev, out, pos, time_emb, bt_cls)
    D = x.shape[-1]
    assert x.is_contiguous() and x.dtype == torch.float16 and x.numel() == B * T * (P + 1) * D
    if out is None:
        out = torch.empty(B + B * P * T, D, dtype=torch.float16, device=x.device)
    assert out.is_contiguous() and out.shape == (B + B * P * T, D)
    if z_prev is not None:
        assert z_prev.is_contiguous() and z_prev.shape == out.shape
    else:
        assert pos.is_contiguous() and pos.shape == (P + 1, D) and time_emb.is_contiguous() and time_emb.shape[0] >= T
        assert bt_cls.numel() == D
    rc = _lib.load().stllm_bt_vit_to_branch_f16(_ptr(x), _ptr(z_prev), _ptr(out), _ptr(pos), _ptr(time_emb),
                                                _ptr(bt_cls), B, T, P, D, _stream())
    _lib.check(rc, "stllm_bt_vit_to_branch_f16")
    launch_count += 1
    return out


def bt_branch_to_vit_avg(x, z, B, T, P, out=None):
    """(x + branch rows back in ViT order) / 2, x [(B*T)*(P+1), D], z [B + B*P*T, D]"""
    global launch_count
    _need_cuda(x, z, out)
    D = x.shape[-1]
    assert x.is_contiguous() and z.is_contiguous() and z.shape == (B + B * P * T, D)
    if out is None:
        out = torch.empty_like(x)
    rc = _lib.load().stllm_bt_branch_to_vit_avg_f16(_ptr(x), _ptr(z), _ptr(out), B, T, P, D, _stream())
    _lib.check(rc, "stllm_bt_branch_to_vit_avg_f16")
    launch_count += 1
    return out


def cosine_loss(a, b):
    """mean over rows of 2 - 2 cos(a_row, b_row); a, b [rows, C] fp16 -> fp32 scalar tensor"""
    global launch_count
    _need_cuda(a, b)
    rows, lda = _rows2d(a)
    _, ldb = _rows2d(b)
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    rc = _lib.load().stllm_cosine_loss_f16(_ptr(a), lda, _ptr(b), ldb, rows, a.shape[-1], _ptr(out), _stream())
    _lib.check(rc, "stllm_cosine_loss_f16")
    launch_count += 1
    return out[0]


def cross_entropy(logits, labels, ignore_index=-100):
    """mean CE over rows with label != ignore_index; logits [rows, V] fp16, labels int64 [rows]"""
    global launch_count
    _need_cuda(logits, labels)
    rows, ldl = _rows2d(logits)
    labels = labels.to(torch.int64).contiguous().view(-1)
    assert labels.numel() == rows
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    rc = _lib.load().stllm_cross_entropy_f16(_ptr(logits), ldl, _ptr(labels), rows, logits.shape[-1], ignore_index,
                                             _ptr(out), _stream())
    _lib.check(rc, "stllm_cross_entropy_f16")
    launch_count += 1
    return out[0] / out[1]


def rope_tables(head_dim, max_pos, base=10000.0, device="cuda"):
    """fp32 cos/sin tables [max_pos, head_dim/2] (modeling_llama_mem.py:81-110 inv_freq/outer/cos/sin)"""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    return freqs.cos().contiguous().to(device), freqs.sin().contiguous().to(device)


# ---------------------------------------------------------------------------------------------------------------------
# single-token decode step (csrc/decode.cu): HBM-bound byte movers for R <= 8 rows (beams)
# ---------------------------------------------------------------------------------------------------------------------
DECODE_MAX_ROWS = 8


def decode_linear(x, w, out, norm_mode=NORM_NONE, norm_weight=None, norm_eps=0.0, residual=None, epilogue=EPI_NONE):
    """out[R, N'] = epi(norm(x)[R, K] @ w[N, K].T) (+ residual);  R <= 8.  norm_mode NORM_RMS: `w` carries the norm
    weight (packed copy), the result is scaled by rstd;  2: explicit RMSNorm with `norm_weight` (reference rounding).
    out may alias residual."""
    global launch_count
    _need_cuda(x, w, out, norm_weight, residual)
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and out.dtype == torch.float16
    assert x.dim() == 2 and out.dim() == 2 and x.stride(1) == 1 and out.stride(1) == 1 and w.stride(1) == 1
    R, K = x.shape
    N = w.shape[0]
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    assert w.shape[1] == K and out.shape == (R, n_out) and R <= DECODE_MAX_ROWS
    ldr = 0
    if residual is not None:
        assert residual.dtype == torch.float16 and residual.shape == (R, n_out) and residual.stride(1) == 1
        ldr = residual.stride(0)
    if norm_mode == 2:
        assert norm_weight is not None and norm_weight.dtype == torch.float16 and norm_weight.numel() == K
    rc = _lib.load().stllm_decode_linear_f16(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), R,
                                             N, K, norm_mode, _ptr(norm_weight), float(norm_eps), _ptr(residual), ldr,
                                             epilogue, _stream())
    _lib.check(rc, "stllm_decode_linear_f16")
    launch_count += 1
    return out


def decode_rope_append(qkv, cos_t, sin_t, prefix_len, n_gen, k_suffix, v_suffix, src, nheads, head_dim):
    """RoPE at position prefix_len + n_gen[0] on the q and k parts of qkv [R, 3*H*D] (q in place); roped k and v are
    appended to the suffix caches [rows, l_max, H*D] at token n_gen[0]; src[r, n_gen[0]] = r."""
    global launch_count
    _need_cuda(qkv, cos_t, sin_t, n_gen, k_suffix, v_suffix, src)
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[1] == 3 * nheads * head_dim
    assert n_gen.dtype == torch.int32 and src.dtype == torch.int32 and src.dim() == 2 and src.stride(1) == 1
    assert k_suffix.dim() == 3 and k_suffix.stride(2) == 1 and k_suffix.shape == v_suffix.shape
    assert k_suffix.stride() == v_suffix.stride() and k_suffix.shape[2] == nheads * head_dim
    assert cos_t.dtype == torch.float32 and cos_t.shape[1] == head_dim // 2 and cos_t.is_contiguous()
    R = qkv.shape[0]
    assert k_suffix.shape[0] >= R and src.shape[0] >= R and src.shape[1] >= k_suffix.shape[1]
    rc = _lib.load().stllm_decode_rope_append_f16(_ptr(qkv), qkv.stride(0), _ptr(cos_t), _ptr(sin_t), cos_t.shape[0],
                                                  int(prefix_len), _ptr(n_gen), _ptr(k_suffix), _ptr(v_suffix),
                                                  k_suffix.stride(0), k_suffix.stride(1), k_suffix.shape[1], _ptr(src),
                                                  src.stride(0), R, nheads, head_dim, _stream())
    _lib.check(rc, "stllm_decode_rope_append_f16")
    launch_count += 1


def decode_attention_workspace(nheads, n_split, device):
    n = int(_lib.load().stllm_decode_attention_workspace_bytes(nheads, n_split))
    return torch.empty(n // 4, dtype=torch.float32, device=device)


def decode_attention(q, k_prefix, v_prefix, k_suffix, v_suffix, src, n_gen, out, workspace, nheads, head_dim, scale,
                     n_split=4, prefix_mask=None):
    """out[r] = softmax(scale q_r K^T) V over [k_prefix (shared by all beams) || generated tokens 0..n_gen[0] of beam r
    through the ancestry table src].  q: [R, >= H*D] (first H*D columns), k_prefix/v_prefix: [Sp, H*D] (row stride free),
    k_suffix/v_suffix: [rows, l_max, H*D], out: [R, H*D].  Launches 2 kernels."""
    global launch_count
    _need_cuda(q, k_prefix, v_prefix, k_suffix, v_suffix, src, n_gen, out, workspace, prefix_mask)
    R = q.shape[0]
    HD = nheads * head_dim
    assert q.dim() == 2 and q.stride(1) == 1 and q.shape[1] >= HD and out.shape == (R, HD) and out.stride(1) == 1
    assert k_prefix.dim() == 2 and k_prefix.stride(1) == 1 and k_prefix.shape[1] == HD
    assert v_prefix.shape == k_prefix.shape and v_prefix.stride() == k_prefix.stride()
    assert k_suffix.dim() == 3 and k_suffix.stride(2) == 1 and k_suffix.shape[2] == HD
    assert v_suffix.shape == k_suffix.shape and v_suffix.stride() == k_suffix.stride()
    assert n_gen.dtype == torch.int32 and src.dtype == torch.int32 and src.stride(1) == 1
    if prefix_mask is not None:
        assert prefix_mask.dtype == torch.uint8 and prefix_mask.is_contiguous() and prefix_mask.numel() == k_prefix.shape[0]
    rc = _lib.load().stllm_decode_attention_f16(
        _ptr(q), q.stride(0), _ptr(k_prefix), _ptr(v_prefix), k_prefix.stride(0), _ptr(prefix_mask),
        k_prefix.shape[0], _ptr(k_suffix), _ptr(v_suffix), k_suffix.stride(0), k_suffix.stride(1), k_suffix.shape[1],
        _ptr(src), src.stride(0), _ptr(n_gen), _ptr(out), out.stride(0), _ptr(workspace),
        workspace.numel() * workspace.element_size(), R, nheads, head_dim, n_split, float(scale), _stream())
    _lib.check(rc, "stllm_decode_attention_f16")
    launch_count += 2
    return out


def decode_attention_fused_workspace(nheads, n_split, device):
    """zeroed workspace (split partials + per-head tickets) of decode_attention_fused"""
    n = int(_lib.load().stllm_decode_attention_fused_workspace_bytes(nheads, n_split))
    return torch.zeros(n // 4, dtype=torch.float32, device=device)


def decode_attention_fused(qkv, cos_t, sin_t, k_prefix, v_prefix, k_suffix, v_suffix, src, n_gen, out, workspace, nheads,
                           head_dim, scale, n_split=4, prefix_mask=None, prefetch=None, prefetch2=None):
    """decode_rope_append + decode_attention in one launch: qkv [R, 3*H*D] un-roped (left unchanged); the new token's
    roped k and v are appended to the suffix caches at token n_gen[0]; out [R, H*D].  prefetch / prefetch2: optional
    contiguous tensors (weights the next kernels stream) pulled into L2 meanwhile."""
    global launch_count
    _need_cuda(qkv, cos_t, sin_t, k_prefix, v_prefix, k_suffix, v_suffix, src, n_gen, out, workspace, prefix_mask)
    R = qkv.shape[0]
    HD = nheads * head_dim
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[1] == 3 * HD and out.shape == (R, HD) and out.stride(1) == 1
    assert k_prefix.dim() == 2 and k_prefix.stride(1) == 1 and k_prefix.shape[1] == HD
    assert v_prefix.shape == k_prefix.shape and v_prefix.stride() == k_prefix.stride()
    assert k_suffix.dim() == 3 and k_suffix.stride(2) == 1 and k_suffix.shape[2] == HD and k_suffix.shape[0] >= R
    assert v_suffix.shape == k_suffix.shape and v_suffix.stride() == k_suffix.stride()
    assert n_gen.dtype == torch.int32 and src.dtype == torch.int32 and src.stride(1) == 1 and src.shape[0] >= R
    assert cos_t.dtype == torch.float32 and cos_t.shape[1] == head_dim // 2 and cos_t.is_contiguous()
    if prefix_mask is not None:
        assert prefix_mask.dtype == torch.uint8 and prefix_mask.is_contiguous() and prefix_mask.numel() == k_prefix.shape[0]
    rc = _lib.load().stllm_decode_attention_fused_f16(
        _ptr(qkv), qkv.stride(0), _ptr(cos_t), _ptr(sin_t), cos_t.shape[0], _ptr(k_prefix), _ptr(v_prefix),
        k_prefix.stride(0), _ptr(prefix_mask), k_prefix.shape[0], _ptr(k_suffix), _ptr(v_suffix), k_suffix.stride(0),
        k_suffix.stride(1), k_suffix.shape[1], _ptr(src), src.stride(0), _ptr(n_gen), _ptr(out), out.stride(0),
        _ptr(workspace), workspace.numel() * workspace.element_size(), R, nheads, head_dim, n_split, float(scale),
        _ptr(prefetch), prefetch.numel() * prefetch.element_size() if prefetch is not None else 0,
        _ptr(prefetch2), prefetch2.numel() * prefetch2.element_size() if prefetch2 is not None else 0, _stream())
    _lib.check(rc, "stllm_decode_attention_fused_f16")
    launch_count += 1
    return out
