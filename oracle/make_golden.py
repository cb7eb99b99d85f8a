"""Generate tests/golden/*.npz by IMPORTING THE REFERENCE ITSELF (uncbiag/LiftReg).

Run in the build container, where /root/reference exists (it does not travel to
the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Every output tensor below is produced by the reference's own functions from
/root/reference/src (torch 2.11 CPU, numpy 2.3), on seeded synthetic inputs that are
stored alongside.  Two shims are needed to import/run the reference here and are
recorded in the fixtures' `meta`:
  * np.float = float             (removed alias used at sdct:141,182,207; layers.py:169)
  * torch.Tensor.cuda = identity (hard-coded .cuda() at net_utils.py:87; the container has no GPU)
The reference has no tests or golden vectors of its own, so these
fixtures are the pin for oracle/liftreg_oracle.c and oracle/torch_port.py.
"""
import os
import sys

import numpy as np

np.float = float  # shim 1
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self  # shim 2

REF = os.environ.get("LIFTREG_REF", "/root/reference/src")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from liftreg.utils import sdct_projection_utils as sdct  # noqa: E402
from liftreg.utils import net_utils  # noqa: E402
from liftreg.layers import layers as ref_layers  # noqa: E402
from liftreg_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
META = "reference=/root/reference/src torch=%s numpy=%s shims=np.float,Tensor.cuda" % (torch.__version__, np.__version__)
CPU = torch.device("cpu")


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array(META), **arrs)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def small_volume(shape, seed):
    hu = synthetic.ct_phantom(shape, seed=seed, sigma=1.0, noise_hu=5.0, nodules=6)
    return sdct.calc_relative_atten_coef(hu), hu


def g_ray_grid():
    """sdct:15-57 project_grid_multi, even and odd detector sizes, + proj_layer variant (layers.py:194-236)."""
    spacing = torch.tensor([2.2, 1.7, 2.5])
    for tag, (rd, rh), shp in (("even", (14, 12), (10, 18, 12)), ("odd", (15, 9), (11, 13, 17))):
        poses = synthetic.wrapper_poses(60.0, 3, shp[1])
        grid, dx = sdct.project_grid_multi(poses, (rd, rh), [1, 1, 1], shp, spacing, CPU, torch.float32)
        save("ray_grid_" + tag, poses=poses, resolution=np.array([rd, rh]), obj_shape=np.array(shp),
             spacing=spacing.numpy(), grid=grid.numpy(), dx=dx.numpy())
    # proj_layer's own grid builder (y/w variant)
    shp = (10, 18, 12)
    layer = ref_layers.proj_layer(spacing, 1.5, 40.0, 3, shp, (8, 8), CPU)
    save("ray_grid_projlayer", poses=layer.poses_scale * shp[1], obj_shape=np.array(shp), spacing=spacing.numpy(),
         resolution=np.array(layer.grids.shape[1:3]), grid_flipped=layer.grids.numpy(), dx=layer.dx.numpy())


def g_drr():
    """sdct:59-100 calculate_projection on CPU + the pre-sum samples the same call sequence produces."""
    shp = (24, 20, 28)
    mu, hu = small_volume(shp, 7)
    poses = synthetic.wrapper_poses(60.0, 4, shp[1])
    res = (36, 42)
    spacing = (2.2, 2.2, 2.2)
    proj = sdct.calculate_projection(mu, poses, res, [1, 1, 1], spacing, CPU)
    vol = torch.from_numpy(mu)[None, None]
    grid, dx = sdct.project_grid_multi(poses, res, [1, 1, 1], shp, torch.tensor(spacing), CPU, torch.float32)
    samples = F.grid_sample(vol, torch.flip(grid, [4]).reshape(1, 1, 1, -1, 3), align_corners=True).reshape(4, 36, 42, 20)
    save("drr_small", vol=mu, hu=hu, poses=poses, resolution=np.array(res), spacing=np.array(spacing, np.float32),
         proj=proj, samples=samples.numpy())
    # geometry given in a CSV (sdct:161-177) only differs by poses = csv/spacing: exercise arbitrary poses
    poses2 = np.array([[-31.25, 77.5, 3.125], [12.0, 64.0, -9.5]])
    proj2 = sdct.calculate_projection(mu, poses2, (30, 30), [1, 1, 1], (1.0, 2.0, 3.0), CPU)
    save("drr_small_csvposes", vol=mu, poses=poses2, resolution=np.array((30, 30)),
         spacing=np.array((1.0, 2.0, 3.0), np.float32), proj=proj2)


def g_drr_full():
    """cfg 1 (BASELINE.json configs[0]): 160^3, 4 views / 60 deg, 240^2 detector, reference CPU path.
    Stored: a strided subset and float64 checksums (the full image is 0.9 MB)."""
    hu = synthetic.ct_phantom((160, 160, 160))
    mu = sdct.calc_relative_atten_coef(hu)
    poses = synthetic.wrapper_poses(60.0, 4, 160)
    proj = sdct.calculate_projection(mu, poses, (240, 240), [1, 1, 1], (2.2, 2.2, 2.2), CPU)
    save("drr_cfg1", poses=poses, proj_sub=proj[:, ::6, ::6].copy(), proj_row=proj[:, 120, :].copy(),
         sum64=np.array(proj.astype(np.float64).sum()), sumsq64=np.array((proj.astype(np.float64) ** 2).sum()),
         per_view_norm=np.sqrt((proj.astype(np.float64) ** 2).sum(axis=(1, 2))),
         mu_sum64=np.array(mu.astype(np.float64).sum()))
    return mu, poses, proj


def g_proj_layer():
    """layers.py:159-192 proj_layer forward and its autograd gradient wrt x."""
    shp = (12, 16, 14)
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.rand(2, *shp).astype(np.float32)).requires_grad_(True)
    spacing = torch.tensor([2.2, 2.2, 2.2])
    layer = ref_layers.proj_layer(spacing, 1.5, 60.0, 3, shp, (10, 12), CPU)
    y = layer(x)
    gy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
    y.backward(gy)
    save("proj_layer", x=x.detach().numpy(), spacing=spacing.numpy(), out=y.detach().numpy(), grad_out=gy.numpy(),
         grad_x=x.grad.numpy(), in_shape=np.array(shp), out_shape=np.array((10, 12)),
         resolution_scale=np.array(1.5), scan_range=np.array(60.0), proj_num=np.array(3))


def g_backproj():
    """sdct:227-250 grid + LiftRegDeformSubspaceBackproj.py:85-93 sampling block (replayed verbatim)."""
    rs = np.random.RandomState(11)
    shp = (12, 20, 16)
    pshape = (24, 22)
    B, P = 2, 3
    poses = np.repeat(synthetic.wrapper_poses(60.0, P, shp[1])[None], B, 0).astype(np.float32)
    target_proj = torch.from_numpy(rs.uniform(-1, 1, (B, P) + pshape).astype(np.float32)).requires_grad_(True)
    grids = sdct.backproj_grids_with_poses(poses[0:1], shp, pshape, device=CPU)
    perm = grids.permute(0, 1, 3, 4, 5, 2)
    w, d, h = shp
    vol = F.grid_sample(target_proj.reshape(B * P, 1, *pshape),
                        perm.expand(B, -1, -1, -1, -1, -1).reshape(B * P, w * d, h, -1),
                        align_corners=True, padding_mode="zeros").reshape(B, P, w, d, h)
    go = torch.from_numpy(rs.randn(*vol.shape).astype(np.float32))
    vol.backward(go)
    save("backproj_small", poses=poses, img_shape=np.array(shp), proj_shape=np.array(pshape),
         target_proj=target_proj.detach().numpy(), grid=grids.numpy(), out=vol.detach().numpy(),
         grad_out=go.numpy(), grad_proj=target_proj.grad.numpy())
    # older fixed-geometry builder (sdct:179-202): emitter at 3.0*w
    g_old = sdct.backproj_grids(60.0, P, shp, pshape, device=CPU)
    save("backproj_grids_old", grid=g_old.numpy(), img_shape=np.array(shp), proj_shape=np.array(pshape),
         scan_range=np.array(60.0), proj_num=np.array(P))


def g_backproj_full(mu, poses, proj):
    """cfg 2 (configs[1]) backprojection: 4 x 256^2 -> 160^3, inputs = normalised DRRs of the phantom."""
    proj256 = sdct.calculate_projection(mu, poses, (256, 256), [1, 1, 1], (2.2, 2.2, 2.2), CPU)
    tp = torch.from_numpy(synthetic.normalise_projection(proj256))[None]
    grids = sdct.backproj_grids_with_poses(poses[None].astype(np.float32), (160, 160, 160), (256, 256), device=CPU)
    perm = grids.permute(0, 1, 3, 4, 5, 2)
    vol = F.grid_sample(tp.reshape(4, 1, 256, 256), perm.reshape(4, 160 * 160, 160, -1), align_corners=True,
                        padding_mode="zeros").reshape(1, 4, 160, 160, 160).numpy()
    save("backproj_cfg2", poses=poses, proj256_sub=proj256[:, ::8, ::8].copy(),
         proj256_sum64=np.array(proj256.astype(np.float64).sum()),
         out_sub=vol[:, :, ::10, ::10, ::10].copy(), out_line=vol[0, :, 80, 80, :].copy(),
         sum64=np.array(vol.astype(np.float64).sum()),
         per_view_norm=np.sqrt((vol.astype(np.float64) ** 2).sum(axis=(0, 2, 3, 4))))


def g_warp():
    """net_utils.py:9-56 Bilinear in every (zero_boundary, using_scale, mode) combination + autograd grads."""
    rs = np.random.RandomState(5)
    B, C, shp = 2, 2, (9, 12, 10)
    img = rs.uniform(-1, 1, (B, C) + shp).astype(np.float32)
    idm = net_utils.identity_map(shp).numpy()
    disp = synthetic.smooth_displacement(shp, seed=5, max_disp=0.35, coarse=4)
    phi = (idm[None] + np.stack([disp, -disp])).astype(np.float32)   # pushes some samples outside [-1,1]
    arrs = dict(img=img, phi=phi, identity=idm)
    for zb in (False, True):
        for us in (False, True):
            for mode in ("bilinear", "nearest"):
                ti = torch.from_numpy(img).requires_grad_(True)
                tp = torch.from_numpy(phi).requires_grad_(True)
                out = net_utils.Bilinear(zero_boundary=zb, using_scale=us, mode=mode)(ti, tp)
                key = "zb%d_us%d_%s" % (zb, us, mode)
                arrs["out_" + key] = out.detach().numpy()
                if mode == "bilinear":
                    go = torch.from_numpy(np.random.RandomState(6).randn(*out.shape).astype(np.float32))
                    out.backward(go)
                    arrs["grad_out"] = go.numpy()
                    arrs["gimg_" + key] = ti.grad.numpy()
                    arrs["gphi_" + key] = tp.grad.numpy()
    save("warp_small", **arrs)
    save("identity_map", id_160=net_utils.identity_map((160, 160, 160)).numpy()[:, ::16, ::16, ::16].copy(),
         id_7_9_11=net_utils.identity_map((7, 9, 11)).numpy(),
         gen_5_6_7=net_utils.gen_identity_map([5, 6, 7], 1.0).numpy())


def g_warp_full():
    """cfg 2 warp: 160^3, Bilinear(zero_boundary=True, using_scale=True) as the model builds it (model :28)."""
    hu = synthetic.ct_phantom((160, 160, 160))
    moving = synthetic.hu_to_unit(hu)[None, None]
    idm = net_utils.identity_map((160, 160, 160)).numpy()
    phi = (synthetic.smooth_displacement((160, 160, 160)) + idm)[None]
    out = net_utils.Bilinear(zero_boundary=True, using_scale=True)(torch.from_numpy(moving), torch.from_numpy(phi)).numpy()
    save("warp_cfg2", out_sub=out[0, 0, ::8, ::8, ::8].copy(), out_line=out[0, 0, 80, 80, :].copy(),
         sum64=np.array(out.astype(np.float64).sum()), norm64=np.array(np.sqrt((out.astype(np.float64) ** 2).sum())),
         phi_sum64=np.array(phi.astype(np.float64).sum()), moving_sum64=np.array(moving.astype(np.float64).sum()))


def g_atten():
    rs = np.random.RandomState(2)
    hu = rs.uniform(-1500, 1500, (4, 5, 6)).astype(np.float32)
    save("atten", hu=hu, mu=sdct.calc_relative_atten_coef(hu),
         mu_inplace=sdct.calc_relative_atten_coef_cuda(torch.from_numpy(hu.copy())).numpy())


def g_ncc():
    """layers/losses.py:14-29 NCCLoss (value and gradient wrt the input).  The module imports mermaid.finite_differences
    at its top (not installable here, not used by NCCLoss): an empty stand-in module lets the import succeed."""
    import types
    for name in ("mermaid", "mermaid.finite_differences"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mermaid"].finite_differences = sys.modules["mermaid.finite_differences"]
    from liftreg.layers import losses as ref_losses
    rs = np.random.RandomState(77)
    shape = (3, 1, 12, 10, 14)
    target = rs.uniform(-1, 1, shape).astype(np.float32)
    warped = (0.7 * target + 0.3 * rs.uniform(-1, 1, shape)).astype(np.float32)
    x = torch.from_numpy(warped).requires_grad_(True)
    loss = ref_losses.NCCLoss()(x, torch.from_numpy(target))
    loss.backward()
    save("ncc", warped=warped, target=target, loss=loss.detach().numpy(), grad=x.grad.numpy())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1:                      # regenerate selected fixtures only: python oracle/make_golden.py ncc ...
        for name in sys.argv[1:]:
            globals()["g_" + name]()
        sys.exit(0)
    g_ncc()
    g_ray_grid()
    g_drr()
    g_proj_layer()
    g_backproj()
    g_warp()
    g_atten()
    mu, poses, proj = g_drr_full()
    g_backproj_full(mu, poses, proj)
    g_warp_full()
