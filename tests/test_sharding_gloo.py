"""N>1 host logic on CPU: world_size-2 gloo process groups, the oracle injected as the compute function.

Checks that the view-sharded DRR (+ all-gather) and the z-slab-sharded backprojection / warp reassemble to exactly
the unsharded result, including uneven splits (odd view counts / odd plane counts) and ranks with no work."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_fns():
    from oracle import c_oracle

    def project(vol_b, poses, res, spacing, y_mode, out_scale):
        return torch.from_numpy(c_oracle.drr_forward(vol_b.numpy(), poses, res, spacing, y_mode, out_scale))

    def backproject(tp, poses, shape, slab):
        full = c_oracle.backproject_forward(tp.numpy(), poses, shape)
        return torch.from_numpy(np.ascontiguousarray(full[:, :, slab[0]:slab[0] + slab[1]]))

    def warp(img, phi_slab, z0):
        # oracle on the full grid, then cut the slab: pad phi to full depth with the slab in place
        B, _, D, H, W = img.shape
        full_phi = np.zeros((B, 3, D, H, W), np.float32)
        full_phi[:, :, z0:z0 + phi_slab.shape[2]] = phi_slab.numpy()
        out = c_oracle.warp_forward(img.numpy(), full_phi, True, True, "bilinear")
        return torch.from_numpy(np.ascontiguousarray(out[:, :, z0:z0 + phi_slab.shape[2]]))

    return project, backproject, warp


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from liftreg_b200 import sharding, synthetic
        from oracle import c_oracle
        project, backproject, warp = _oracle_fns()
        rs = np.random.RandomState(0)
        res = {}

        # --- view-sharded DRR, 2 items x 3 views = 6 views over `world` ranks (uneven for world=4), + all-gather
        vol = rs.rand(2, 6, 8, 7).astype(np.float32)
        poses = synthetic.wrapper_poses(60.0, 3, 8)
        full = sharding.drr_project_sharded(torch.from_numpy(vol), poses, (9, 10), (1.0, 1.0, 1.0), group=None,
                                            project_fn=project)
        ref = c_oracle.drr_forward(vol, poses, (9, 10), (1.0, 1.0, 1.0))
        res["drr"] = bool(np.array_equal(full.numpy(), ref))
        mine, my_views = sharding.drr_project_sharded(torch.from_numpy(vol), poses, (9, 10), (1.0, 1.0, 1.0),
                                                      gather=False, project_fn=project)
        res["drr_local"] = bool(np.array_equal(mine.numpy(), ref.reshape(6, 9, 10)[my_views]))
        res["drr_range"] = my_views

        # --- single view, more ranks than views: some ranks idle, result still complete
        one = sharding.drr_project_sharded(torch.from_numpy(vol[:1]), poses[:1], (9, 10), (1.0, 1.0, 1.0), project_fn=project)
        res["drr_idle"] = bool(np.array_equal(one.numpy(), ref[:1, :1]))

        # --- z-slab backprojection: d = 7 planes (uneven), gather to check
        tp = rs.uniform(-1, 1, (2, 2, 12, 11)).astype(np.float32)
        bposes = synthetic.wrapper_poses(60.0, 2, 9).astype(np.float32)
        shape = (7, 9, 6)
        slab, (z0, z1) = sharding.backproject_sharded(torch.from_numpy(tp), bposes, shape, backproject_fn=backproject)
        bref = c_oracle.backproject_forward(tp, bposes, shape)
        res["bp_slab"] = bool(np.array_equal(slab.numpy(), bref[:, :, z0:z1]))
        fullv, _ = sharding.backproject_sharded(torch.from_numpy(tp), bposes, shape, gather=True, backproject_fn=backproject)
        res["bp_full"] = bool(np.array_equal(fullv.numpy(), bref))

        # --- z-slab warp: phi sharded like the output, image replicated
        img = rs.uniform(-1, 1, (1, 2) + shape).astype(np.float32)
        phi = (c_oracle.identity_map(shape)[None] + rs.uniform(-0.3, 0.3, (1, 3) + shape)).astype(np.float32)
        phi_slab, zr = sharding.shard_along_z(torch.from_numpy(phi), world, rank)
        wslab = sharding.warp_sharded(torch.from_numpy(img), phi_slab, zr, zero_boundary=True, warp_fn=warp)
        wref = c_oracle.warp_forward(img, phi, True, True, "bilinear")
        res["warp_slab"] = bool(np.array_equal(wslab.numpy(), wref[:, :, zr[0]:zr[1]]))
        wfull = sharding.warp_sharded(torch.from_numpy(img), phi_slab, zr, zero_boundary=True, warp_fn=warp, gather=True)
        res["warp_full"] = bool(np.array_equal(wfull.numpy(), wref))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_ops_reassemble_exactly(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = []
    for r in range(world):
        res = results[r]
        ranges.append(res.pop("drr_range"))
        assert all(res.values()), (r, res)
    # the ranks' view lists (dealt round-robin: view v -> rank v % world) cover [0, 6) exactly once
    assert sorted(v for r in ranges for v in r) == list(range(6))
    assert all(v % world == r for r in range(world) for v in ranges[r])


def test_split_range_properties():
    from liftreg_b200.sharding import all_ranges, split_range
    for n in (0, 1, 4, 7, 64, 160, 161):
        for world in (1, 2, 3, 4, 8):
            rr = all_ranges(n, world)
            assert rr[0][0] == 0 and rr[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
            sizes = [hi - lo for lo, hi in rr]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    assert split_range(64, 8, 3) == (24, 32)          # cfg 4: 64 views over 8 GPUs
    assert split_range(160, 8, 7) == (140, 160)       # cfg 5: 160 planes over 8 GPUs
    with pytest.raises(ValueError):
        split_range(4, 2, 2)
