import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    sd = {}
    rest = {}
    for k in z.files:
        v = z[k]
        if k.startswith("w."):
            t = torch.from_numpy(v)
            sd[k[2:]] = t.float() if t.is_floating_point() else t
        else:
            rest[k] = v
    return sd, rest


def assert_close(got, ref, atol, rtol, what=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu() if isinstance(ref, torch.Tensor) else torch.from_numpy(np.asarray(ref)).float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol)
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} out of tolerance, max_abs_err={err.max():.3e}, "
                           f"ref_max={ref.abs().max():.3e}")
