"""Functional PyTorch surface over the C-ABI: tensors in, tensors out, autograd wired.

PyTorch is plumbing here (device memory, streams, autograd graph); all arithmetic happens in
libliftreg_b200.so.  CUDA float32 tensors only -- there is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _native

PAD_ZEROS, PAD_BORDER = 0, 1
MODE_LINEAR, MODE_NEAREST = 0, 1
YNORM_WM1, YNORM_W = 0, 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _need_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: liftreg_b200 has no CPU path (got device %s)" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    return t.contiguous()


def _poses64(poses):
    """(P,3) or (n_sets,P,3) -> contiguous float64 numpy (n_sets,P,3)."""
    if isinstance(poses, torch.Tensor):
        poses = poses.detach().cpu().numpy()
    a = np.ascontiguousarray(poses, dtype=np.float64)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("poses must have shape (P,3) or (B,P,3), got %s" % (a.shape,))
    return a


def _poses32(poses):
    if isinstance(poses, torch.Tensor):
        poses = poses.detach().cpu().numpy()
    a = np.ascontiguousarray(poses, dtype=np.float32)
    if a.ndim == 3:
        a = a[0]            # geometry frozen from batch item 0 (reference model :85-87)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("poses must have shape (P,3) or (B,P,3), got %s" % (a.shape,))
    return np.ascontiguousarray(a)


def _spacing3(spacing):
    if isinstance(spacing, torch.Tensor):
        spacing = spacing.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(spacing, dtype=np.float32).reshape(-1))
    if a.size != 3:
        raise ValueError("spacing must have 3 entries")
    return a


def _fp(a):
    return a.ctypes.data_as(_native.c_float_p)


def _dp(a):
    return a.ctypes.data_as(_native.c_double_p)


# --------------------------------------------------------------------------- DRR
class _DRR(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, poses, rd, rh, spacing, y_norm_mode, out_scale, out):
        vol = _need_cuda_f32(vol, "vol")
        B, d, w, h = vol.shape
        n_sets, P, _ = poses.shape
        if out is None:
            proj = torch.empty((B, P, rd, rh), device=vol.device, dtype=torch.float32)
        else:
            if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == B * P * rd * rh):
                raise ValueError("out must be a contiguous CUDA float32 tensor of %d elements" % (B * P * rd * rh))
            proj = out
            ctx.mark_dirty(out)
        with torch.cuda.device(vol.device):
            _native.check(_native.lib().lr_drr_forward(_ptr(vol), B, d, w, h, _dp(poses), n_sets, P, rd, rh,
                                                       _fp(spacing), y_norm_mode, out_scale, _ptr(proj), _stream()),
                          "lr_drr_forward")
        ctx.geom = (poses, rd, rh, spacing, y_norm_mode, out_scale, (B, d, w, h))
        return proj

    @staticmethod
    def backward(ctx, grad_proj):
        poses, rd, rh, spacing, y_norm_mode, out_scale, (B, d, w, h) = ctx.geom
        grad_proj = _need_cuda_f32(grad_proj, "grad_proj")
        n_sets, P, _ = poses.shape
        grad_vol = torch.zeros((B, d, w, h), device=grad_proj.device, dtype=torch.float32)
        with torch.cuda.device(grad_proj.device):
            _native.check(_native.lib().lr_drr_backward(_ptr(grad_proj), B, d, w, h, _dp(poses), n_sets, P, rd, rh,
                                                        _fp(spacing), y_norm_mode, out_scale, _ptr(grad_vol), _stream()),
                          "lr_drr_backward")
        return grad_vol, None, None, None, None, None, None, None


def drr_project(vol, poses, resolution, spacing, y_norm_mode=YNORM_WM1, out_scale=0.1, out=None):
    """Cone-beam DRR of vol (B,d,w,h) -> (B,P,rd,rh); differentiable wrt vol.

    Replaces reference sdct:59-86 (y_norm_mode=0, out_scale=0.1) and layers.py:182-187 (y_norm_mode=1, out_scale=1).
    poses: (P,3) shared by the batch, or (B,P,3); float64, voxel units.
    out: optional pre-allocated contiguous (B,P,rd,rh) CUDA tensor (or any contiguous view with that many elements,
    e.g. a rank's slot of an all-gather buffer) that the kernel writes into; it is returned.
    """
    poses = _poses64(poses)
    if vol.dim() != 4:
        raise ValueError("vol must be (B,d,w,h), got %s" % (tuple(vol.shape),))
    if poses.shape[0] not in (1, vol.shape[0]):
        raise ValueError("poses batch (%d) must be 1 or B (%d)" % (poses.shape[0], vol.shape[0]))
    rd, rh = int(resolution[0]), int(resolution[1])
    return _DRR.apply(vol, poses, rd, rh, _spacing3(spacing), int(y_norm_mode), float(out_scale), out)


def drr_project_peers(vol, poses, resolution, spacing, out_ptrs, view_stride, y_norm_mode=YNORM_WM1, out_scale=0.1):
    """Multi-GPU sweep form of drr_project (no autograd): the kernel stores the images into every buffer of `out_ptrs`
    (raw device addresses: this rank's gather buffer and the peers', see sharding.PeerGather), view k of the call at
    k * view_stride images from each address.  Nothing is returned; the caller orders the ranks with a barrier."""
    import ctypes
    poses = _poses64(poses)
    vol = _need_cuda_f32(vol, "vol")
    if vol.dim() != 4 or poses.shape[0] not in (1, vol.shape[0]):
        raise ValueError("vol must be (B,d,w,h) and poses (P,3) or (B,P,3)")
    B, d, w, h = vol.shape
    n_sets, P, _ = poses.shape
    ptrs = (ctypes.c_void_p * len(out_ptrs))(*[int(a) for a in out_ptrs])
    with torch.cuda.device(vol.device):
        _native.check(_native.lib().lr_drr_forward_peers(_ptr(vol), B, d, w, h, _dp(poses), n_sets, P, int(resolution[0]),
                                                         int(resolution[1]), _fp(_spacing3(spacing)), int(y_norm_mode),
                                                         float(out_scale), ptrs, len(out_ptrs), int(view_stride), _stream()),
                      "lr_drr_forward_peers")


def project_grid(poses, resolution, obj_shape, spacing, device, y_norm_mode=YNORM_WM1, flip=False, want_grid=True):
    """Materialised sample grid (P,rd,rh,w,3) and dx (P,rd,rh) -- reference sdct:15-57, for API parity only."""
    poses = _poses64(poses)[0]
    P = poses.shape[0]
    rd, rh = int(resolution[0]), int(resolution[1])
    d, w, h = (int(s) for s in obj_shape)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("project_grid needs a CUDA device: liftreg_b200 has no CPU path")
    grid = torch.empty((P, rd, rh, w, 3), device=dev, dtype=torch.float32) if want_grid else None
    dx = torch.empty((P, rd, rh), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _native.check(_native.lib().lr_project_grid(_dp(poses), P, rd, rh, d, w, h, _fp(_spacing3(spacing)),
                                                    int(y_norm_mode), int(bool(flip)), _ptr(grid), _ptr(dx), _stream()),
                      "lr_project_grid")
    return grid, dx


# --------------------------------------------------------------------------- backprojection
_PLAN_CACHE = {}          # geometry -> device plan buffer (insertion-ordered; a handful of geometries per process)
_PLAN_CACHE_MAX = 8


def backproject_plan(poses, proj_shape, img_shape, device):
    """Device buffer holding the geometry plan of lr_backproject_forward_planned for (poses, proj_shape, img_shape),
    built once per geometry and cached -- the counterpart of the reference caching its sample grid on the first
    batch (LiftRegDeformSubspaceBackproj.py:85-87).  poses: (P,3) float32.  (Measured slower than rebuilding the
    tables in shared memory -- backproject() does not use it; kept as a tested entry point.)"""
    poses = _poses32(poses)
    P = poses.shape[0]
    pw, ph = (int(s) for s in proj_shape)
    d, w, h = (int(s) for s in img_shape)
    dev = torch.device(device)
    key = (poses.tobytes(), P, pw, ph, d, w, h, dev.index if dev.index is not None else torch.cuda.current_device())
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        lib = _native.lib()
        nbytes = lib.lr_backproject_plan_bytes(P, pw, ph, d, w, h)
        if nbytes == 0:
            raise ValueError("bad backprojection geometry %r" % ((P, pw, ph, d, w, h),))
        plan = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _native.check(lib.lr_backproject_plan_build(_fp(poses), P, pw, ph, d, w, h, _ptr(plan), nbytes, _stream()),
                          "lr_backproject_plan_build")
            torch.cuda.current_stream().synchronize()      # once per geometry: later calls may come from any stream
        while len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        _PLAN_CACHE[key] = plan
    return plan


class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj, poses, d_total, w, h, out, channel_offset, i_begin, i_count):
        proj = _need_cuda_f32(proj, "target_proj")
        B, P, pw, ph = proj.shape
        d = i_count
        nv = d * w * h
        if out is None:
            out = torch.empty((B, P, d, w, h), device=proj.device, dtype=torch.float32)
            view, bs, cs = out, P * nv, nv
        else:
            if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()):
                raise ValueError("out must be a contiguous CUDA float32 tensor")
            if out.dim() != 5 or out.shape[0] != B or tuple(out.shape[2:]) != (d, w, h) \
                    or out.shape[1] < channel_offset + P:
                raise ValueError("out must be (B,>=%d,%d,%d,%d), got %s" % (channel_offset + P, d, w, h, tuple(out.shape)))
            view, bs, cs = out[:, channel_offset:channel_offset + P], out.shape[1] * nv, nv
            ctx.mark_dirty(out)
        with torch.cuda.device(proj.device):
            _native.check(_native.lib().lr_backproject_forward_slab(_ptr(proj), _fp(poses), B, P, pw, ph, d_total, w, h,
                                                                    i_begin, i_count, ctypes.c_void_p(view.data_ptr()),
                                                                    bs, cs, _stream()),
                          "lr_backproject_forward_slab")
        ctx.geom = (poses, B, P, pw, ph, d, w, h, channel_offset, out.shape[1], d_total, i_begin)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        poses, B, P, pw, ph, d, w, h, off, nchan, d_total, i_begin = ctx.geom
        if d != d_total:
            raise NotImplementedError("gradient wrt the projections is not implemented for slab-sharded backprojection "
                                      "(the reference detaches this output, LiftRegDeformSubspaceBackproj.py:93)")
        grad_out = _need_cuda_f32(grad_out, "grad_out")
        nv = d * w * h
        gview = grad_out[:, off:off + P]
        grad_proj = torch.zeros((B, P, pw, ph), device=grad_out.device, dtype=torch.float32)
        with torch.cuda.device(grad_out.device):
            _native.check(_native.lib().lr_backproject_backward(ctypes.c_void_p(gview.data_ptr()), nchan * nv, nv, _fp(poses),
                                                                B, P, pw, ph, d, w, h, _ptr(grad_proj), _stream()),
                          "lr_backproject_backward")
        # `out` was overwritten in channels [off, off+P): the gradient wrt its previous content passes through elsewhere
        grad_buf = None
        if ctx.needs_input_grad[5]:
            grad_buf = grad_out.clone()
            grad_buf[:, off:off + P] = 0
        return grad_proj, None, None, None, None, grad_buf, None, None, None


def backproject(target_proj, poses, img_shape, out=None, channel_offset=0, slab=None):
    """Lift projections (B,P,pw,ph) into a volume (B,P,d,w,h); differentiable wrt target_proj.

    Replaces reference sdct:227-250 + LiftRegDeformSubspaceBackproj.py:85-93 in one kernel (no 131 MB grid).
    poses: (P,3) or (B,P,3) float32 (item 0 is used, as the reference freezes geometry from the first batch).
    out/channel_offset: optionally write into channels [offset, offset+P) of a pre-allocated (B,Ctot,d,w,h)
    buffer, which removes the torch.cat of model :95-98; the whole buffer is returned.
    slab=(i_begin, i_count): compute only planes [i_begin, i_begin+i_count) of axis 0 (multi-GPU z-slab sharding);
    the result (and `out`) then has i_count planes.
    """
    if target_proj.dim() != 4:
        raise ValueError("target_proj must be (B,P,pw,ph), got %s" % (tuple(target_proj.shape),))
    poses = _poses32(poses)
    if poses.shape[0] != target_proj.shape[1]:
        raise ValueError("poses has %d views but target_proj has %d" % (poses.shape[0], target_proj.shape[1]))
    d, w, h = (int(s) for s in img_shape)
    i_begin, i_count = (0, d) if slab is None else (int(slab[0]), int(slab[1]))
    if not (0 <= i_begin and i_count > 0 and i_begin + i_count <= d):
        raise ValueError("slab %r is not inside [0, %d)" % (slab, d))
    return _Backproject.apply(target_proj, poses, d, w, h, out, int(channel_offset), i_begin, i_count)


def backproj_grid(poses, img_shape, proj_shape, device):
    """Materialised voxel->detector grid (P,2,d,w,h) -- reference sdct:227-250, for API parity only."""
    poses = _poses32(poses)
    d, w, h = (int(s) for s in img_shape)
    pw, ph = (int(s) for s in proj_shape)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("backproj_grid needs a CUDA device: liftreg_b200 has no CPU path")
    grid = torch.empty((poses.shape[0], 2, d, w, h), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _native.check(_native.lib().lr_backproj_grid(_fp(poses), poses.shape[0], d, w, h, pw, ph, _ptr(grid), _stream()),
                      "lr_backproj_grid")
    return grid


# --------------------------------------------------------------------------- warp
class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, phi, padding, mode, using_scale, disp_plus_identity, z_begin):
        img = _need_cuda_f32(img, "input1")
        phi = _need_cuda_f32(phi, "input2")
        B, C, D, H, W = img.shape
        z_count = phi.shape[2]
        out = torch.empty((B, C, z_count, H, W), device=img.device, dtype=torch.float32)
        with torch.cuda.device(img.device):
            _native.check(_native.lib().lr_warp_forward_slab(_ptr(img), _ptr(phi), B, C, D, H, W, z_begin, z_count, padding,
                                                             mode, int(using_scale), int(disp_plus_identity), _ptr(out),
                                                             _stream()),
                          "lr_warp_forward_slab")
        ctx.save_for_backward(img, phi)
        ctx.cfg = (padding, mode, using_scale, disp_plus_identity, z_begin)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        img, phi = ctx.saved_tensors
        padding, mode, using_scale, ident, z_begin = ctx.cfg
        grad_out = _need_cuda_f32(grad_out, "grad_out")
        B, C, D, H, W = img.shape
        z_count = phi.shape[2]
        need_img, need_phi = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_img and mode == MODE_NEAREST:
            raise NotImplementedError("gradient wrt the image is not implemented for nearest-mode warps")
        gimg = torch.zeros_like(img) if need_img else None
        gphi = torch.empty_like(phi) if need_phi else None
        with torch.cuda.device(img.device):
            _native.check(_native.lib().lr_warp_backward_slab(_ptr(grad_out), _ptr(img), _ptr(phi), B, C, D, H, W, z_begin,
                                                              z_count, padding, mode, int(using_scale), int(ident),
                                                              _ptr(gimg), _ptr(gphi), _stream()),
                          "lr_warp_backward_slab")
        return gimg, gphi, None, None, None, None, None


def warp(img, phi, zero_boundary=False, using_scale=True, mode="bilinear", disp_plus_identity=False, z_begin=None):
    """Spatial transformer: img (B,C,D,H,W) sampled at phi (B,3,D,H,W) in [-1,1]; differentiable wrt both.

    Replaces reference net_utils.py:26-56.  disp_plus_identity=True treats phi as a displacement and adds the
    identity map in-kernel (fuses LiftRegDeformSubspaceBackproj.py:68).
    z_begin: phi is a z-slab (B,3,Dz,H,W) holding output planes [z_begin, z_begin+Dz) of the full (D,H,W) image
    (multi-GPU z-slab sharding); the result is the matching (B,C,Dz,H,W) slab.
    """
    if mode not in ("bilinear", "nearest"):
        raise ValueError("mode must be 'bilinear' or 'nearest', got %r" % (mode,))
    if img.dim() != 5 or phi.dim() != 5 or phi.shape[1] != 3 or phi.shape[0] != img.shape[0] \
            or tuple(phi.shape[3:]) != tuple(img.shape[3:]):
        raise ValueError("expected input1 (B,C,D,H,W) and input2 (B,3,D,H,W); got %s and %s"
                         % (tuple(img.shape), tuple(phi.shape)))
    if z_begin is None:
        if phi.shape[2] != img.shape[2]:
            raise ValueError("input2 has %d planes but input1 has %d (pass z_begin for a slab)" % (phi.shape[2], img.shape[2]))
        z_begin = 0
    if not (0 <= z_begin and z_begin + phi.shape[2] <= img.shape[2]):
        raise ValueError("slab [%d, %d) is not inside [0, %d)" % (z_begin, z_begin + phi.shape[2], img.shape[2]))
    return _Warp.apply(img, phi, PAD_ZEROS if zero_boundary else PAD_BORDER,
                       MODE_LINEAR if mode == "bilinear" else MODE_NEAREST, bool(using_scale), bool(disp_plus_identity),
                       int(z_begin))


def identity_map(sz, device):
    """Normalised identity map (3,*sz) on `device` -- reference net_utils.py:59-87."""
    D, H, W = (int(s) for s in sz)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("identity_map needs a CUDA device: liftreg_b200 has no CPU path")
    out = torch.empty((3, D, H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _native.check(_native.lib().lr_identity_map(D, H, W, _ptr(out), _stream()), "lr_identity_map")
    return out


def atten_coef_(img):
    """In-place HU -> attenuation on a CUDA tensor -- reference sdct:11-13."""
    t = _need_cuda_f32(img, "img")
    if t.data_ptr() != img.data_ptr():
        raise ValueError("atten_coef_ needs a contiguous tensor")
    with torch.cuda.device(t.device):
        _native.check(_native.lib().lr_atten_coef(_ptr(t), t.numel(), _ptr(t), _stream()), "lr_atten_coef")
    return img


# --------------------------------------------------------------------------- PCA-subspace decode (row f2)
class _PcaDecode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coefs, basis, mean, D, H, W, add_identity):
        coefs = _need_cuda_f32(coefs, "coefs")
        basis = _need_cuda_f32(basis, "pca_vectors")
        mean_c = _need_cuda_f32(mean, "pca_mean") if mean is not None else None
        B, K = coefs.shape
        N = basis.shape[0]
        out = torch.empty((B, N), device=coefs.device, dtype=torch.float32)
        with torch.cuda.device(coefs.device):
            _native.check(_native.lib().lr_pca_decode(_ptr(coefs), _ptr(basis), _ptr(mean_c), B, K, N, int(add_identity),
                                                      D, H, W, _ptr(out), _stream()), "lr_pca_decode")
        ctx.save_for_backward(basis)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (basis,) = ctx.saved_tensors
        # d/dcoefs = grad_out (B,N) @ basis (N,K): the second pass over the basis; basis and mean are frozen buffers
        grad_out = _need_cuda_f32(grad_out, "grad_out")
        B, N = grad_out.shape
        K = basis.shape[1]
        gcoefs = torch.zeros((B, K), device=grad_out.device, dtype=torch.float32)
        with torch.cuda.device(grad_out.device):
            _native.check(_native.lib().lr_pca_decode_backward(_ptr(grad_out), _ptr(basis), B, K, N, _ptr(gcoefs), _stream()),
                          "lr_pca_decode_backward")
        return gcoefs, None, None, None, None, None, None


def pca_decode(coefs, pca_vectors, pca_mean=None, img_shape=None, add_identity=False):
    """disp = F.linear(coefs, pca_vectors, pca_mean) as one streaming kernel (reference model :102); with
    add_identity (needs img_shape = (D,H,W), N = 3*D*H*W) the identity map of model :68 is added too, so the result
    reshaped to (B,3,D,H,W) is the map the warp consumes.  coefs (B,K); pca_vectors (N,K) as the model stores it
    (:42); pca_mean (N).  Differentiable wrt coefs."""
    if coefs.dim() != 2 or pca_vectors.dim() != 2 or coefs.shape[1] != pca_vectors.shape[1]:
        raise ValueError("expected coefs (B,K) and pca_vectors (N,K); got %s and %s" % (tuple(coefs.shape), tuple(pca_vectors.shape)))
    if pca_mean is not None and tuple(pca_mean.shape) != (pca_vectors.shape[0],):
        raise ValueError("pca_mean must be (N,)")
    if not pca_vectors.is_contiguous():
        # a silent .contiguous() here would transpose-copy the whole basis (2.75 GB at 160^3) on EVERY call
        raise ValueError("pca_vectors must be a dense row-major (N,K) tensor; the reference builds an (N,K) view of a "
                         "(K,N) array (model :42) -- call .contiguous() on it once (dropin.install() does)")
    D = H = W = 0
    if add_identity:
        if img_shape is None or 3 * int(np.prod(img_shape)) != pca_vectors.shape[0]:
            raise ValueError("add_identity needs img_shape with 3*D*H*W == N")
        D, H, W = (int(s) for s in img_shape)
    out = _PcaDecode.apply(coefs, pca_vectors, pca_mean, D, H, W, bool(add_identity))
    if img_shape is not None:
        return out.reshape(coefs.shape[0], 3, *[int(s) for s in img_shape])
    return out


# --------------------------------------------------------------------------- NCC similarity (row f4)
class _Ncc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x = _need_cuda_f32(x, "input")
        y = _need_cuda_f32(y, "target")
        B = x.shape[0]
        N = x.numel() // B
        sums = torch.empty((B, 7), device=x.device, dtype=torch.float64)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().lr_ncc_sums(_ptr(x), _ptr(y), B, N, _ptr(sums), _stream()), "lr_ncc_sums")
        ncc = sums[:, 2] / torch.sqrt(sums[:, 3] * sums[:, 4])           # the 1/N of the three means cancels
        ctx.save_for_backward(x, y, sums)
        return (1.0 - ncc.mean()).to(torch.float32)

    @staticmethod
    def backward(ctx, grad_loss):
        x, y, sums = ctx.saved_tensors
        B = x.shape[0]
        N = x.numel() // B
        g = grad_loss.to(device=x.device, dtype=torch.float32).reshape(1).contiguous()
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().lr_ncc_backward(_ptr(x), _ptr(y), B, N, _ptr(sums), _ptr(g), _ptr(gx), _stream()),
                          "lr_ncc_backward")
        return gx, None


def ncc_loss(input, target):
    """1 - mean over the batch of the normalised cross correlation -- reference layers/losses.py:14-29 (NCCLoss).
    Differentiable wrt `input` (the warped image); `target` is data."""
    if input.shape != target.shape:
        raise ValueError("input %s and target %s must have the same shape" % (tuple(input.shape), tuple(target.shape)))
    return _Ncc.apply(input, target)


# --------------------------------------------------------------------------- displacement regulariser (row f4)
FD_BOUNDARY = {"linear": 0, "neumann_zero": 1}       # LR_FD_LINEAR / LR_FD_NEUMANN_ZERO (mermaid FD modes)


class _DiffusionReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, boundary):
        disp = _need_cuda_f32(disp, "disp")
        B, _, D, H, W = disp.shape
        total = torch.empty(1, device=disp.device, dtype=torch.float64)
        with torch.cuda.device(disp.device):
            _native.check(_native.lib().lr_diffusion_reg_sum(_ptr(disp), B, D, H, W, boundary, _ptr(total), _stream()),
                          "lr_diffusion_reg_sum")
        ctx.save_for_backward(disp)
        ctx.boundary = boundary
        return (total[0] / float(B * D * H * W)).to(torch.float32)

    @staticmethod
    def backward(ctx, grad_loss):
        disp, = ctx.saved_tensors
        B, _, D, H, W = disp.shape
        g = grad_loss.to(device=disp.device, dtype=torch.float32).reshape(1).contiguous()
        gd = torch.empty_like(disp)
        with torch.cuda.device(disp.device):
            _native.check(_native.lib().lr_diffusion_reg_backward(_ptr(disp), B, D, H, W, ctx.boundary, _ptr(g), _ptr(gd),
                                                                  _stream()), "lr_diffusion_reg_backward")
        return gd, None


def diffusion_reg(disp, boundary="linear"):
    """mean over (B, D, H, W) of the nine squared central differences of a displacement field (B,3,D,H,W) -- reference
    losses/SubspaceLoss.py:51-67 (compute_reg_loss), one fused pass instead of ~27 kernels; differentiable.

    `boundary` names the rule of mermaid's finite differences on the volume faces: "linear" (FD_torch's default mode:
    the missing neighbour is extrapolated linearly) or "neumann_zero" (zero difference on the faces)."""
    if disp.dim() != 5 or disp.shape[1] != 3:
        raise ValueError("disp must be (B,3,D,H,W), got %s" % (tuple(disp.shape),))
    if boundary not in FD_BOUNDARY:
        raise ValueError("boundary must be one of %s, got %r" % (sorted(FD_BOUNDARY), boundary))
    return _DiffusionReg.apply(disp, FD_BOUNDARY[boundary])
