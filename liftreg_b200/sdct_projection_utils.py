"""Drop-in mirror of the reference's src/liftreg/utils/sdct_projection_utils.py ("sdct").

Same function names, argument meaning and return types, so main.py / eval.py / tools/preprocessingDRR.py work
unchanged when this module stands in for the reference one (see liftreg_b200.dropin).  Every function that
touched torch.nn.functional.grid_sample, or built a sample grid, now calls one fused sm_100a kernel through the
C-ABI.  The misspelt names (`calculate_projection_wraper`) are the reference's.

Differences a caller can observe:
  * only sample_rate [1,1,1] (the only value the reference ever passes, sdct:152,171,218,253) and float32;
  * a CUDA device is mandatory (the reference hard-codes "cuda" in the wrappers, sdct:154,173);
  * np.float (removed in numpy 1.24; sdct:141,182,207) is not used, so the wrappers run on current numpy.
Additional fused entry points: backproject() and DRRProjector (device-resident volumes, no per-call H2D).
"""
import ctypes

import numpy as np
import torch
from numpy import genfromtxt

from . import _native, ops


# ------------------------------------------------------------------ HU -> attenuation (sdct:6-13)
def calc_relative_atten_coef(img):
    """numpy HU -> linear attenuation, water = 0.2 (sdct:6-9). Host-side numpy, as in the reference."""
    new_img = img.astype(np.float32).copy()
    new_img[new_img < -1000] = -1000
    return (new_img + 1000.) / 1000. * 0.2


def calc_relative_atten_coef_cuda(img):
    """Tensor HU -> attenuation; clamps `img` in place like the reference (sdct:11-13) and returns a new tensor."""
    if img.is_cuda and img.dtype == torch.float32 and img.is_contiguous():
        img[img < -1000] = -1000
        out = img.clone()
        return ops.atten_coef_(out)
    img[img < -1000] = -1000
    return (img + 1000.) / 1000. * 0.2


# ------------------------------------------------------------------ helpers
def _check_sample_rate(sample_rate):
    if [int(s) for s in sample_rate] != [1, 1, 1]:
        raise NotImplementedError("only sample_rate [1,1,1] is supported (the reference never uses another value)")


def _cuda_device(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("liftreg_b200 runs on CUDA devices only (got %s); there is no CPU path" % (dev,))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _wrapper_poses_scale(scan_range, proj_num, emitter_y):
    """sdct:139-144 / :180-185 / :205-210 pose synthesis (float64, in units of the coronal size)."""
    angle_half = scan_range / 2.
    poses_scale = np.ndarray((proj_num, 3), dtype=np.float64)
    poses_scale[:, 1] = emitter_y
    poses_scale[:, 0] = np.tan(np.linspace(-angle_half, angle_half, num=proj_num) / 180. * np.pi) * 3.
    poses_scale[:, 2] = np.linspace(-0.2, 0.2, num=proj_num)
    return poses_scale


def _default_resolution(shape, receptor_size):
    if receptor_size is not None:
        return list(receptor_size)
    resolution_scale = 1.5                                   # sdct:149-151
    return [int(shape[0] * resolution_scale), int(shape[2] * resolution_scale)]


# ------------------------------------------------------------------ ray grids (sdct:15-57)
def project_grid_multi(emi_pos, resolution, sample_rate, obj_shape, spacing, device, dtype):
    """(grid (P,rd,rh,w,3), dx (P,rd,rh)) on `device` -- same tensors as sdct:15-57, bit for bit.
    Kept for API parity; the DRR kernels never materialise the grid."""
    _check_sample_rate(sample_rate)
    if dtype not in (torch.float32, torch.float):
        raise NotImplementedError("project_grid_multi: float32 only")
    return ops.project_grid(emi_pos, resolution, obj_shape, spacing, _cuda_device(device), ops.YNORM_WM1, flip=False)


# ------------------------------------------------------------------ DRR (sdct:59-100)
class DRRProjector:
    """Reusable DRR context: owns a pinned host staging area and a device workspace so that repeated
    calculate_projection calls (tools/preprocessingDRR.py:123-148 does two per case) do not re-allocate.
    `project_numpy` is calculate_projection's numpy-in/numpy-out contract through lr_drr_forward_host."""

    def __init__(self, device="cuda"):
        self.device = _cuda_device(device)
        self._ws = None
        self._pin_in = None
        self._pin_out = None

    def _buffers(self, n_in, n_out, ws_bytes):
        if self._ws is None or self._ws.numel() < ws_bytes:
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        if self._pin_in is None or self._pin_in.numel() < n_in:
            self._pin_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
        if self._pin_out is None or self._pin_out.numel() < n_out:
            self._pin_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
        return self._ws, self._pin_in, self._pin_out

    def project_numpy(self, img, poses, resolution, spacing, y_norm_mode=ops.YNORM_WM1, out_scale=0.1):
        img = np.ascontiguousarray(img, dtype=np.float32)
        squeeze = img.ndim == 3
        if squeeze:
            img = img[None]
        B, d, w, h = img.shape
        poses64 = ops._poses64(poses)
        n_sets, P, _ = poses64.shape
        rd, rh = int(resolution[0]), int(resolution[1])
        sp = ops._spacing3(spacing)
        lib = _native.lib()
        ws_bytes = lib.lr_drr_forward_host_workspace_bytes(B, d, w, h, P, rd, rh)
        n_out = B * P * rd * rh
        ws, pin_in, pin_out = self._buffers(img.size, n_out, ws_bytes)
        pin_in[:img.size].copy_(torch.from_numpy(img.reshape(-1)))
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream()
            _native.check(lib.lr_drr_forward_host(ctypes.c_void_p(pin_in.data_ptr()), B, d, w, h, ops._dp(poses64), n_sets, P,
                                                  rd, rh, ops._fp(sp), int(y_norm_mode), float(out_scale),
                                                  ctypes.c_void_p(pin_out.data_ptr()), ctypes.c_void_p(ws.data_ptr()),
                                                  ws.numel(), ctypes.c_void_p(st.cuda_stream)),
                          "lr_drr_forward_host")
        out = pin_out[:n_out].numpy().reshape(B, P, rd, rh).copy()
        return out[0] if squeeze else out


_default_projector = {}


def _projector(device):
    dev = _cuda_device(device)
    key = (dev.type, dev.index)
    if key not in _default_projector:
        _default_projector[key] = DRRProjector(dev)
    return _default_projector[key]


def calculate_projection(img, poses, resolution, sample_rate, spacing, device):
    """numpy (d,w,h) attenuation -> numpy (P,rd,rh) float32 DRR, synchronous (sdct:59-100).

    Geometry (sdct:61-68): detector = XZ plane through the origin, Y axis towards the emitter, poses in voxels.
    One fused kernel replaces grid build + flip + grid_sample + sum + *dx + *0.1."""
    _check_sample_rate(sample_rate)
    return _projector(device).project_numpy(img, poses, resolution, spacing, ops.YNORM_WM1, 0.1)


def calculate_projection_wraper(img_3d, scan_range, proj_num, spacing, receptor_size=None):
    """(proj (P,rd,rh), poses (P,3) float64) for an arc of `scan_range` degrees (sdct:138-159)."""
    poses_scale = _wrapper_poses_scale(scan_range, proj_num, 3.5)
    resolution = _default_resolution(img_3d.shape, receptor_size)
    sample_rate = [int(1), int(1), int(1)]
    device = torch.device("cuda")                            # sdct:154
    poses = poses_scale * img_3d.shape[1]
    img_proj = calculate_projection(img_3d, poses, resolution, sample_rate, spacing, device)
    return img_proj, poses


def calculate_projection_wraper_with_geo_csv_file(img_3d, img_spacing, geo_path, receptor_size=None):
    """Same with emitter positions (mm) read from a CSV with one header line (sdct:161-177)."""
    geo_txt = genfromtxt(geo_path, delimiter=',')[1:]
    poses = geo_txt / img_spacing
    resolution = _default_resolution(img_3d.shape, receptor_size)
    sample_rate = [int(1), int(1), int(1)]
    device = torch.device("cuda")                            # sdct:173
    img_proj = calculate_projection(img_3d, poses, resolution, sample_rate, img_spacing, device)
    return img_proj, poses


# ------------------------------------------------------------------ grids for the differentiable paths
def forward_grids(scan_range, proj_num, spacing, img_shape, device=torch.device("cuda"), receptor_size=None):
    """(grids flipped to grid_sample order (P,rd,rh,w,3), dx) with the emitter at 3.0*w (sdct:204-225)."""
    poses = _wrapper_poses_scale(scan_range, proj_num, 3.) * img_shape[1]
    resolution = _default_resolution(img_shape, receptor_size)
    return ops.project_grid(poses, resolution, img_shape, spacing, _cuda_device(device), ops.YNORM_WM1, flip=True)


def forward_grids_with_poses(poses, spacing, img_shape, device=torch.device("cuda"), receptor_size=None):
    """Same for explicit poses (sdct:252-265)."""
    resolution = _default_resolution(img_shape, receptor_size)
    return ops.project_grid(poses, resolution, img_shape, spacing, _cuda_device(device), ops.YNORM_WM1, flip=True)


def backproj_grids(scan_range, proj_num, img_shape, proj_shape, device=torch.device("cuda")):
    """Voxel->detector grid (P,2,d,w,h) for the synthetic arc with the emitter at 3.0*w (sdct:179-202).
    NB the reference computes this variant as scale*x + trans rather than (x-s)*scale+s; the two differ in the
    last ulp, and this mirror evaluates the (x-s)*scale+s form of sdct:227-250."""
    poses = (_wrapper_poses_scale(scan_range, proj_num, 3.) * img_shape[1]).astype(np.float32)
    return ops.backproj_grid(poses, img_shape, proj_shape, _cuda_device(device))


def backproj_grids_with_poses(poses, img_shape, proj_shape, device=torch.device("cuda")):
    """Voxel->detector grid (B,P,2,d,w,h) (sdct:227-250); B pose sets as in the reference."""
    poses = np.asarray(poses)
    if poses.ndim != 3:
        raise ValueError("poses must be (B,P,3)")
    dev = _cuda_device(device)
    return torch.stack([ops.backproj_grid(poses[b], img_shape, proj_shape, dev) for b in range(poses.shape[0])])


# ------------------------------------------------------------------ fused backprojection (new entry point)
def backproject(target_proj, poses, img_shape, out=None, channel_offset=0):
    """target_proj (B,P,pw,ph) -> (B,P,d,w,h): sdct:227-250 + LiftRegDeformSubspaceBackproj.py:85-93 in one kernel."""
    return ops.backproject(target_proj, poses, img_shape, out=out, channel_offset=channel_offset)
