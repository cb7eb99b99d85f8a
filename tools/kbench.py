#!/usr/bin/env python
"""Quick per-kernel timing loop for kernel development (not the contract benchmark -- that is bench.py).

    python tools/kbench.py [warp] [backproject] [drr] [--iters N]

Times each kernel at the cfg2 / cfg1 sizes with rotating buffers (cold L2) under a CUDA graph and prints us per launch."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liftreg_b200 import _native, ops, synthetic  # noqa: E402

VOL, DET, P, R = (160, 160, 160), (256, 256), 4, 8


def main():
    iters = 2000
    args = list(sys.argv[1:])
    if "--iters" in args:
        k = args.index("--iters")
        iters = int(args[k + 1])
        del args[k:k + 2]
    batch = 1
    if "--batch" in args:        # batch items per launch for warp / backproject (BASELINE configs[2] uses 8)
        k = args.index("--batch")
        batch = int(args[k + 1])
        del args[k:k + 2]
    which = args or ["warp", "backproject", "drr"]
    dev = torch.device("cuda:0")
    lib = _native.lib()
    stream = torch.cuda.Stream()
    rs = np.random.RandomState(0)
    nv = VOL[0] * VOL[1] * VOL[2]
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    poses = synthetic.wrapper_poses(60.0, P, VOL[1])
    poses32 = np.ascontiguousarray(poses.astype(np.float32))
    phi = torch.from_numpy((synthetic.smooth_displacement(VOL) + synthetic.identity_map_np(VOL))[None]).to(dev).repeat(batch, 1, 1, 1, 1)
    moving = torch.from_numpy(rs.uniform(-1, 1, (1, 1) + VOL).astype(np.float32)).to(dev).repeat(batch, 1, 1, 1, 1)
    proj = torch.from_numpy(rs.uniform(-1, 1, (1, P) + DET).astype(np.float32)).to(dev).repeat(batch, 1, 1, 1)
    mu = torch.from_numpy(rs.uniform(0, 0.3, (1,) + VOL).astype(np.float32)).to(dev)
    sets = [dict(phi=phi.clone(), moving=moving.clone(), proj=proj.clone(), mu=mu.clone(),
                 warped=torch.empty((batch, 1) + VOL, device=dev), lifted=torch.empty((batch, P) + VOL, device=dev),
                 drr=torch.empty((1, P, 240, 240), device=dev), drr256=torch.empty((1, P, 256, 256), device=dev)) for _ in range(R)]
    sp3 = np.array([2.2, 2.2, 2.2], np.float32)
    p64 = np.ascontiguousarray(poses, np.float64)

    def k_warp(s, st):
        _native.check(lib.lr_warp_forward(vp(s["moving"]), vp(s["phi"]), batch, 1, *VOL, 0, 0, 1, 0, vp(s["warped"]), st), "warp")

    def k_backproject(s, st):
        _native.check(lib.lr_backproject_forward(vp(s["proj"]), ops._fp(poses32), batch, P, DET[0], DET[1], *VOL, vp(s["lifted"]),
                                                 P * nv, nv, st), "backproject")

    bp_plan = ops.backproject_plan(poses32, DET, VOL, dev)

    def k_backproject_planned(s, st):
        _native.check(lib.lr_backproject_forward_planned(vp(s["proj"]), vp(bp_plan), batch, P, DET[0], DET[1], *VOL, 0, VOL[0], vp(s["lifted"]),
                                                         P * nv, nv, st), "backproject_planned")

    def k_drr(s, st):
        _native.check(lib.lr_drr_forward(vp(s["mu"]), 1, *VOL, ops._dp(p64), 1, P, 240, 240, ops._fp(sp3), 0,
                                         ctypes.c_float(0.1), vp(s["drr"]), st), "drr")

    def k_drr256(s, st):
        _native.check(lib.lr_drr_forward(vp(s["mu"]), 1, *VOL, ops._dp(p64), 1, P, 256, 256, ops._fp(sp3), 0,
                                         ctypes.c_float(0.1), vp(s["drr256"]), st), "drr256")

    gout = torch.from_numpy(rs.standard_normal((1, 1) + VOL).astype(np.float32)).to(dev)
    gphi = [torch.empty((1, 3) + VOL, device=dev) for _ in range(R)]
    gdrr = torch.from_numpy(rs.standard_normal((1, P, 240, 240)).astype(np.float32)).to(dev)
    gvol = [torch.zeros((1,) + VOL, device=dev) for _ in range(R)]
    for i, s_ in enumerate(sets):
        s_["gphi"], s_["gvol"] = gphi[i], gvol[i]

    def k_warp_bwd(s, st):     # d/dphi only (what training needs: LiftRegDeformSubspaceBackproj.py:69 with moving as data)
        _native.check(lib.lr_warp_backward(vp(gout), vp(s["moving"]), vp(s["phi"]), 1, 1, *VOL, 0, 0, 1, 0, None, vp(s["gphi"]), st), "warp_bwd")

    def k_drr_bwd(s, st):
        _native.check(lib.lr_drr_backward(vp(gdrr), 1, *VOL, ops._dp(p64), 1, P, 240, 240, ops._fp(sp3), 0,
                                          ctypes.c_float(0.1), vp(s["gvol"]), st), "drr_bwd")

    if "pca" in which or "pca_bwd" in which:
        K = 56
        basis = torch.empty((3 * nv, K), device=dev).normal_(0, 1e-3)       # 2.75 GB, as the model holds it
        pmean = torch.zeros(3 * nv, device=dev)
        coefs = torch.randn(1, K, device=dev)
        pout = [torch.empty((1, 3 * nv), device=dev) for _ in range(R)]
        for i, s_ in enumerate(sets):
            s_["pout"] = pout[i]

    if "pca_bwd" in which:
        pgout = torch.randn(1, 3 * nv, device=dev)
        pgc = torch.zeros(1, 56, device=dev)

    def k_pca_bwd(s, st):
        _native.check(lib.lr_pca_decode_backward(vp(pgout), vp(basis), 1, 56, 3 * nv, vp(pgc), st), "pca_bwd")

    def k_pca(s, st):
        _native.check(lib.lr_pca_decode(vp(coefs), vp(basis), vp(pmean), 1, K, 3 * nv, 1, *VOL, vp(s["pout"]), st), "pca")

    ncc_sums = torch.zeros((1, 7), device=dev, dtype=torch.float64)
    ncc_g = torch.ones(1, device=dev)

    def k_ncc(s, st):          # forward of the similarity loss: one pass over warped + target (+ memset and a B-thread finalisation)
        _native.check(lib.lr_ncc_sums(vp(s["warped"]), vp(s["moving"]), 1, nv, vp(ncc_sums), st), "ncc")

    def k_ncc_bwd(s, st):
        _native.check(lib.lr_ncc_backward(vp(s["warped"]), vp(s["moving"]), 1, nv, vp(ncc_sums), vp(ncc_g), vp(s["gvol"]), st), "ncc_bwd")

    reg_sum = torch.zeros(1, device=dev, dtype=torch.float64)

    def k_reg(s, st):          # displacement regulariser (SubspaceLoss.py:51-67): one pass over the 3-channel field
        _native.check(lib.lr_diffusion_reg_sum(vp(s["phi"]), 1, *VOL, 0, vp(reg_sum), st), "reg")

    def k_reg_bwd(s, st):
        _native.check(lib.lr_diffusion_reg_backward(vp(s["phi"]), 1, *VOL, 0, vp(ncc_g), vp(s["gphi"]), st), "reg_bwd")

    def k_atten(s, st):        # HU -> attenuation (sdct:6-13), elementwise
        _native.check(lib.lr_atten_coef(vp(s["moving"]), nv, vp(s["warped"]), st), "atten")

    def k_idmap(s, st):        # identity map (net_utils.py:59-87), write-only
        _native.check(lib.lr_identity_map(*VOL, vp(s["gphi"]), st), "idmap")

    units = {"atten": (k_atten, nv, 8 * nv), "idmap": (k_idmap, nv, 12 * nv), "reg": (k_reg, nv, 12 * nv), "reg_bwd": (k_reg_bwd, nv, 24 * nv), "ncc": (k_ncc, nv, 8 * nv), "ncc_bwd": (k_ncc_bwd, nv, 12 * nv), "warp": (k_warp, batch * nv, batch * 20 * nv), "pca": (k_pca, 3 * nv, 4 * 3 * nv * 56 + 8 * 3 * nv),
             "pca_bwd": (k_pca_bwd, 3 * nv, 4 * 3 * nv * 56 + 4 * 3 * nv), "warp_bwd": (k_warp_bwd, nv, 32 * nv),
             "drr_bwd": (k_drr_bwd, P * 240 * 240 * VOL[1], 4 * nv + 4 * P * 240 * 240), "backproject": (k_backproject, batch * P * nv, batch * (4 * P * nv + 4 * P * DET[0] * DET[1])),
             "backproject_planned": (k_backproject_planned, batch * P * nv, batch * (4 * P * nv + 4 * P * DET[0] * DET[1])),
             "drr": (k_drr, P * 240 * 240 * VOL[1], 4 * nv + 4 * P * 240 * 240),
             "drr256": (k_drr256, P * 256 * 256 * VOL[1], 4 * nv + 4 * P * 256 * 256)}
    for name in which:
        fn, n_units, nbytes = units[name]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            st = ctypes.c_void_p(stream.cuda_stream)
            fn(sets[0], st); stream.synchronize()
            with torch.cuda.graph(g, stream=stream):
                for r in range(R):
                    fn(sets[r], st)
            for _ in range(3):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters // R):
                g.replay()
            e1.record(stream)
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / (iters // R * R)
        print("%-19s %8.2f us  %8.1f G units/s  %7.1f GB/s (algorithmic)" % (name, us, n_units / us * 1e-3, nbytes / us * 1e-3))


if __name__ == "__main__":
    main()
