"""Frame preprocessing of the inference path on the GPU: decoded RGB frames -> encoder input.

Replaces `Chat.transform` of the reference (stllm/conversation/conversation.py:192-200: GroupScale(224, BICUBIC),
GroupCenterCrop(224), Stack, ToTorchFormatTensor, GroupNormalize) + the `view(bt//3, 3, w, h)` of upload_video
(:276-278) -- a per-frame PIL pipeline with a Python loop over channels on the CPU -- by two CUDA kernels
(csrc/preprocess.cu) that reproduce it BIT-EXACTLY: Pillow's 8-bit two-pass bicubic resampler is integer arithmetic
(22-bit fixed-point windows, rounding + clamp to uint8 after each pass), the normalisation three IEEE fp32 operations.
Only the window tables are computed on the host, in double precision, exactly as Pillow's precompute_coeffs does
(src/libImaging/Resample.c); they are cached per frame size.  SURVEY.md 8f rank 3.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib, ops

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)      # conversation.py:192
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)      # conversation.py:193


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_windows(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (bicubic, full-image box): for every output index the
    first input index, the tap count and the 22-bit fixed-point weights.  in_size == out_size gives identity windows
    (Pillow skips the pass; weight 1 << 22 reproduces the input byte exactly)."""
    if in_size == out_size:
        bounds = np.stack((np.arange(out_size), np.ones(out_size, dtype=np.int64)), 1).astype(np.int32)
        return bounds, np.full((out_size, 1), 1 << PRECISION_BITS, dtype=np.int32)
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            v = w * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


class FramePreprocessor:
    """frames -> [n, 3, size, size] fp16 on `device`, the tensor `Chat.upload_video` feeds to encode_img.

    `frames`: a uint8 tensor / array [n, H, W, 3] (RGB, as decord / PIL give them) or a list of PIL images / arrays of
    one size.  Host frames are copied to the device once (n*H*W*3 bytes); everything else runs on the GPU."""

    def __init__(self, size=224, mean=CLIP_MEAN, std=CLIP_STD, device="cuda"):
        self.size = size
        self.mean = (ctypes.c_float * 3)(*mean)
        self.std = (ctypes.c_float * 3)(*std)
        self.device = torch.device(device)
        self._tables = {}

    def _plan(self, H, W):
        key = (H, W)
        if key not in self._tables:
            S = self.size
            # torchvision Resize(int): shorter edge -> S, longer edge = int(S * long / short)
            short, long = (W, H) if W <= H else (H, W)
            new_short, new_long = S, int(S * long / short)
            new_w, new_h = (new_short, new_long) if W <= H else (new_long, new_short)
            assert new_w >= S and new_h >= S
            top, left = int(round((new_h - S) / 2.0)), int(round((new_w - S) / 2.0))       # torchvision center_crop
            xb, xk = resample_windows(W, new_w)
            yb, yk = resample_windows(H, new_h)
            y_first = int(yb[top, 0])
            y_last = int((yb[top:top + S, 0] + yb[top:top + S, 1]).max())
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            self._tables[key] = dict(xb=dev(xb), xk=dev(xk), yb=dev(yb), yk=dev(yk), kx=xk.shape[1], ky=yk.shape[1],
                                     top=top, left=left, y_first=y_first, y_last=y_last)
        return self._tables[key]

    def __call__(self, frames):
        if isinstance(frames, (list, tuple)):
            frames = np.stack([np.asarray(f) for f in frames])
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3, \
            "expected RGB uint8 frames [n, H, W, 3]"
        frames = frames.to(self.device, non_blocking=True).contiguous()
        n, H, W, _ = frames.shape
        t = self._plan(H, W)
        S = self.size
        rows = t["y_last"] - t["y_first"]
        tmp = torch.empty(n * rows * S * 3, dtype=torch.uint8, device=self.device)
        out = torch.empty(n, 3, S, S, dtype=torch.float16, device=self.device)
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        rc = _lib.load().stllm_preprocess_frames_u8(p(frames), n, H, W, S, p(t["xb"]), p(t["xk"]), t["kx"], t["left"],
                                                    p(t["yb"]), p(t["yk"]), t["ky"], t["top"], t["y_first"],
                                                    t["y_last"], ctypes.cast(self.mean, ctypes.c_void_p),
                                                    ctypes.cast(self.std, ctypes.c_void_p), p(tmp), tmp.numel(),
                                                    p(out), ops._stream())
        _lib.check(rc, "stllm_preprocess_frames_u8")
        ops.launch_count += 2
        return out
