"""ctypes front-end of the C oracle (oracle/liftreg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by liftreg_b200/.

numpy in, numpy out; every function cites the reference lines it restates in
liftreg_oracle.c.  Build with `make -C oracle` (or `oracle.c_oracle.build()`).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libliftreg_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile the C restatement with the committed Makefile (gcc, no external deps)."""
    src = os.path.join(_HERE, "liftreg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def set_blend(mode):
    """'aten' (default: the reference's operation order) or 'fast' (the CUDA library's LR_NUMERICS_FAST blend order:
    same indices and weights, blends as fused lerps).  Returns the previous mode."""
    prev = "fast" if lib().lro_get_blend() else "aten"
    lib().lro_set_blend({"aten": 0, "fast": 1}[mode])
    return prev


class blend:
    """with c_oracle.blend('fast'): ..."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = set_blend(self.mode)

    def __exit__(self, *exc):
        set_blend(self.prev)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _poses64(poses):
    return np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)


def grid_sample_3d(vol, grid, padding=0, mode=0):
    vol = _f32(vol); grid = _f32(grid).reshape(-1, 3)
    D, H, W = vol.shape
    out = np.empty(grid.shape[0], np.float32)
    lib().lro_grid_sample_3d(_p(vol), D, H, W, _p(grid), ctypes.c_int64(grid.shape[0]), padding, mode, _p(out))
    return out


def grid_sample_2d(img, grid):
    img = _f32(img); grid = _f32(grid).reshape(-1, 2)
    H, W = img.shape
    out = np.empty(grid.shape[0], np.float32)
    lib().lro_grid_sample_2d(_p(img), H, W, _p(grid), ctypes.c_int64(grid.shape[0]), _p(out))
    return out


def project_grid(poses, resolution, obj_shape, spacing, y_mode=0, want_grid=True):
    """sdct:15-57 -> (grid (P,rd,rh,w,3) in reference (pre-flip) order, dx (P,rd,rh))."""
    poses = _poses64(poses); P = poses.shape[0]
    rd, rh = map(int, resolution); d, w, h = map(int, obj_shape)
    sp = _f32(spacing)
    grid = np.empty((P, rd, rh, w, 3), np.float32) if want_grid else None
    dx = np.empty((P, rd, rh), np.float32)
    lib().lro_project_grid(poses.ctypes.data_as(_f64p), P, rd, rh, d, w, h, _p(sp), y_mode, _p(grid), _p(dx))
    return grid, dx


DRR_SEGS = 4    # liftreg_b200/csrc/drr.cu DRR_SEGS: the kernel sums each ray in 4 runs of ceil(w/4) planes


def kernel_seg_len(w):
    """seg_len argument of drr_forward that reproduces the CUDA kernel's ray-sum order in the current blend mode:
    exact numerics: DRR_SEGS fixed runs of ceil(w / DRR_SEGS) planes; fast numerics (set_blend('fast')): DRR_SEGS equal
    runs over each ray pair's clipped range (negative value = number of balanced runs)."""
    if lib().lro_get_blend():
        return -DRR_SEGS
    return (int(w) + DRR_SEGS - 1) // DRR_SEGS


def drr_forward(vol, poses, resolution, spacing, y_mode=0, out_scale=0.1, acc64=False, want_samples=False, seg_len=0):
    """sdct:59-86.  vol (d,w,h) or (B,d,w,h) -> proj (P,rd,rh) or (B,P,rd,rh).
    seg_len=0: plain sequential ray sum; seg_len=kernel_seg_len(w): the CUDA kernel's ray-segment order."""
    vol = _f32(vol); squeeze = vol.ndim == 3
    if squeeze:
        vol = vol[None]
    B, d, w, h = vol.shape
    poses = _poses64(poses); P = poses.shape[0]
    rd, rh = map(int, resolution)
    sp = _f32(spacing)
    proj = np.empty((B, P, rd, rh), np.float32)
    samples = np.empty((P, rd, rh, w), np.float32) if want_samples else None
    lib().lro_drr_forward(_p(vol), B, d, w, h, poses.ctypes.data_as(_f64p), P, rd, rh, _p(sp), y_mode,
                          ctypes.c_float(out_scale), int(acc64), int(seg_len), _p(proj), _p(samples))
    proj = proj[0] if squeeze else proj
    return (proj, samples) if want_samples else proj


def drr_backward(grad_proj, vol_shape, poses, spacing, y_mode=0, out_scale=0.1):
    """Adjoint wrt the volume.  grad_proj (B,P,rd,rh) -> grad_vol (B,d,w,h)."""
    g = _f32(grad_proj); B, P, rd, rh = g.shape
    d, w, h = map(int, vol_shape)
    poses = _poses64(poses); sp = _f32(spacing)
    gv = np.zeros((B, d, w, h), np.float32)
    lib().lro_drr_backward(_p(g), B, d, w, h, poses.ctypes.data_as(_f64p), P, rd, rh, _p(sp), y_mode,
                           ctypes.c_float(out_scale), _p(gv))
    return gv


def backproj_grid(poses, img_shape, proj_shape):
    """sdct:227-250 for one pose set (P,3) fp32 -> (P,2,d,w,h)."""
    poses = _f32(poses).reshape(-1, 3); P = poses.shape[0]
    d, w, h = map(int, img_shape); pw, ph = map(int, proj_shape)
    grid = np.empty((P, 2, d, w, h), np.float32)
    lib().lro_backproj_grid(_p(poses), P, d, w, h, pw, ph, _p(grid))
    return grid


def backproject_forward(proj, poses, img_shape):
    """LiftRegDeformSubspaceBackproj.py:85-93.  proj (B,P,pw,ph), poses (P,3) -> (B,P,d,w,h)."""
    proj = _f32(proj); B, P, pw, ph = proj.shape
    poses = _f32(poses).reshape(-1, 3); assert poses.shape[0] == P
    d, w, h = map(int, img_shape)
    out = np.empty((B, P, d, w, h), np.float32)
    lib().lro_backproject_forward(_p(proj), _p(poses), B, P, pw, ph, d, w, h, _p(out))
    return out


def backproject_backward(grad_out, poses, proj_shape):
    g = _f32(grad_out); B, P, d, w, h = g.shape
    poses = _f32(poses).reshape(-1, 3)
    pw, ph = map(int, proj_shape)
    gp = np.zeros((B, P, pw, ph), np.float32)
    lib().lro_backproject_backward(_p(g), _p(poses), B, P, pw, ph, d, w, h, _p(gp))
    return gp


def warp_forward(img, phi, zero_boundary=False, using_scale=True, mode="bilinear"):
    """net_utils.py:9-56 Bilinear(zero_boundary, using_scale, mode)(img, phi)."""
    img = _f32(img); phi = _f32(phi)
    B, C, D, H, W = img.shape
    assert phi.shape == (B, 3, D, H, W)
    out = np.empty_like(img)
    lib().lro_warp_forward(_p(img), _p(phi), B, C, D, H, W, 0 if zero_boundary else 1,
                           0 if mode == "bilinear" else 1, int(using_scale), _p(out))
    return out


def warp_backward(grad_out, img, phi, zero_boundary=False, using_scale=True, mode="bilinear",
                  want_img=True, want_phi=True):
    g = _f32(grad_out); img = _f32(img); phi = _f32(phi)
    B, C, D, H, W = img.shape
    gi = np.zeros_like(img) if want_img else None
    gp = np.zeros_like(phi) if want_phi else None
    lib().lro_warp_backward(_p(g), _p(img), _p(phi), B, C, D, H, W, 0 if zero_boundary else 1,
                            0 if mode == "bilinear" else 1, int(using_scale), _p(gi), _p(gp))
    return gi, gp


def identity_map(sz):
    D, H, W = map(int, sz)
    out = np.empty((3, D, H, W), np.float32)
    lib().lro_identity_map(D, H, W, _p(out))
    return out


def atten_coef(hu):
    hu = _f32(hu)
    mu = np.empty_like(hu)
    lib().lro_atten_coef(_p(hu), ctypes.c_int64(hu.size), _p(mu))
    return mu


def pca_decode(coefs, basis, mean=None, img_shape=None):
    """LiftRegDeformSubspaceBackproj.py:102 (+ :68 when img_shape is given): (B,K),(N,K),(N) -> (B,N)."""
    coefs = _f32(coefs); basis = _f32(basis)
    B, K = coefs.shape
    N = basis.shape[0]
    mean = _f32(mean) if mean is not None else None
    D, H, W = (int(s) for s in img_shape) if img_shape is not None else (0, 0, 0)
    out = np.empty((B, N), np.float32)
    lib().lro_pca_decode(_p(coefs), _p(basis), _p(mean), B, K, ctypes.c_int64(N), int(img_shape is not None), D, H, W, _p(out))
    return out
