"""Label / image warping entry point the reference takes from mermaid.

`networks/RegistrationNet.py:191-196` warps the label maps with
    mermaid.utils.compute_warped_image_multiNC(I0, phi, spacing, spline_order=0, zero_boundary=True, use_01_input=...)
mermaid (0.3.2, requirements.txt:61) is not installable here, so this mirror is written from mermaid's published
behaviour: a spatial transformer over the ATen grid sampler (align_corners=True) with nearest (spline_order 0) or linear
(spline_order 1) interpolation, zeros or border padding, map channels in volume-axis order, NO intensity rescaling.
For maps in [-1, 1] (use_01_input=False, what LiftReg's models produce) that is exactly net_utils.Bilinear with
using_scale=False, whose nearest / linear arithmetic is pinned bit-exactly by the goldens.  For use_01_input=True the
map is first rescaled as mermaid's scale_map does ((phi / (spacing*(sz-1)) - 0.5) * 2 per axis); that expression is
restated from the published source and is NOT pinned against a running mermaid ("parity unpinned" for that branch)."""
import numpy as np
import torch

from . import ops


def scale_map(phi, spacing):
    """[0, spacing*(sz-1)] physical coordinates -> [-1, 1] (mermaid map_scale_utils.scale_map)."""
    sz = phi.shape[2:]
    out = torch.empty_like(phi)
    for d in range(len(sz)):
        out[:, d] = (phi[:, d] / (float(spacing[d]) * (sz[d] - 1)) - 0.5) * 2.
    return out


def compute_warped_image_multiNC(I0, phi, spacing, spline_order, zero_boundary=False, use_01_input=True):
    """I0 (B,C,X,Y,Z) sampled at phi (B,3,X,Y,Z); spline_order 0 = nearest (label maps), 1 = trilinear."""
    if spline_order not in (0, 1):
        raise ValueError("only spline_order 0 (nearest) and 1 (linear) are supported")
    if I0.dim() != 5 or phi.dim() != 5:
        raise ValueError("3-D images only: I0 (B,C,X,Y,Z), phi (B,3,X,Y,Z)")
    spacing = np.asarray(spacing, dtype=np.float64).reshape(-1)
    if use_01_input:
        phi = scale_map(phi, spacing)
    I0 = I0 if I0.dtype == torch.float32 else I0.float()
    return ops.warp(I0, phi, zero_boundary=bool(zero_boundary), using_scale=False,
                    mode="nearest" if spline_order == 0 else "bilinear")
