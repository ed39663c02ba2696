"""CPU oracle for the frame preprocessing of the inference path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference turns decoded video frames (PIL RGB images) into the encoder input with
    GroupScale(224, BICUBIC) -> GroupCenterCrop(224) -> Stack -> ToTorchFormatTensor -> GroupNormalize(mean, std)
(stllm/conversation/conversation.py:192-200,274-278; stllm/test/video_transforms.py:66-71,94-123,367-406).  The
arithmetic of the first step lives in third-party code that is not under /root/reference:
    Pillow (requirement.txt pins no version; 12.2.0 installed here), src/libImaging/Resample.c -- the 8-bit two-pass
    (horizontal, then vertical) resampler: double-precision filter weights normalised per output pixel, converted to
    22-bit fixed point, int32 accumulation with round-half-up and clipping to uint8 after EACH pass;
    torchvision.transforms.Resize / CenterCrop (size arithmetic only: torchvision/transforms/functional.py
    _compute_resized_output_size and center_crop).
This file restates that published algorithm in numpy.  Integer path: the uint8 result is BIT-EXACT against Pillow, and
the normalised fp32/fp16 tensor is bit-exact against the reference's own transform classes (pinned by
tests/test_preprocess_oracle.py, which runs the reference's video_transforms.py where /root/reference exists and
otherwise the committed fixtures tests/golden/preprocess_*.npz generated from it).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are 22-bit fixed point for 8-bit channels
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)      # conversation.py:192
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)      # conversation.py:193


def bicubic_filter(x):
    """Resample.c bicubic_filter, a = -0.5"""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size, support=2.0, filt=bicubic_filter):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box (in0 = 0, in1 = in_size).
    Returns bounds int32 [out_size, 2] (first input index, tap count) and kk int32 [out_size, ksize]."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size     # (double)(in1 - in0) / outSize, box is float
    filterscale = scale if scale >= 1.0 else 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            v = w * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w < 0 else int(0.5 + v)       # C (int) cast truncates toward zero
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resized_size(h, w, size):
    """torchvision _compute_resized_output_size for an int size: shorter edge -> size, longer = int(size * long / short)"""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)      # (new_h, new_w)


def crop_offsets(h, w, size):
    """torchvision center_crop: top = int(round((h - size) / 2.0)), left likewise (Python banker's rounding)"""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def _resample_axis(img, bounds, kk, axis):
    """one pass of the 8-bit resampler along `axis` (0: vertical, 1: horizontal) of img [H, W, C] uint8"""
    out_size = bounds.shape[0]
    src = img.astype(np.int64)
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.empty(shape, dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.take(0, axis=axis).shape, 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src.take(xmin + x, axis=axis) * int(kk[xx, x])
        val = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)       # clip8: arithmetic shift, then clamp
        if axis == 0:
            out[xx] = val
        else:
            out[:, xx] = val
    return out


def resize_bicubic_u8(img, size=224):
    """PIL Image.resize((new_w, new_h), BICUBIC) of an RGB uint8 image [H, W, 3] as called by
    torchvision.transforms.Resize(size) (video_transforms.py:119-123): horizontal pass, then vertical pass"""
    h, w = img.shape[:2]
    new_h, new_w = resized_size(h, w, size)
    out = img
    if new_w != w:
        bx, kx = precompute_coeffs(w, new_w)
        out = _resample_axis(out, bx, kx, axis=1)
    if new_h != h:
        by, ky = precompute_coeffs(h, new_h)
        out = _resample_axis(out, by, ky, axis=0)
    return out


def preprocess_frames(frames, size=224, mean=CLIP_MEAN, std=CLIP_STD):
    """frames: iterable of RGB uint8 arrays [H, W, 3] -> float32 [n, 3, size, size], the value sequence of
    Chat.transform + the view(bt//3, 3, w, h) of upload_video (conversation.py:194-200,276-278): resize, centre crop,
    x / 255 (fp32), then (x - mean) / std in fp32 (GroupNormalize: t.sub_(m).div_(s))"""
    outs = []
    for f in frames:
        r = resize_bicubic_u8(np.asarray(f), size)
        top, left = crop_offsets(r.shape[0], r.shape[1], size)
        c = r[top:top + size, left:left + size]
        x = c.astype(np.float32) / np.float32(255)
        x = np.transpose(x, (2, 0, 1)).copy()
        for ch in range(3):
            x[ch] = (x[ch] - np.float32(mean[ch])) / np.float32(std[ch])
        outs.append(x)
    return np.stack(outs)
