"""2-GPU NCCL checks (run under gpurun --gpus 2; skipped when fewer than 2 devices): the sharded entry points over
NCCL reproduce the single-GPU result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from liftreg_b200 import ops, sharding, synthetic
        dev = torch.device("cuda", rank)
        rs = np.random.RandomState(0)
        vol = torch.from_numpy(rs.rand(2, 24, 20, 28).astype(np.float32)).to(dev)
        poses = synthetic.wrapper_poses(60.0, 5, 20)
        full = sharding.drr_project_sharded(vol, poses, (36, 40), (2.2, 2.2, 2.2))        # view-sharded + all-gather
        ref = ops.drr_project(vol, poses, (36, 40), (2.2, 2.2, 2.2))
        ok_drr = bool(torch.equal(full, ref))
        # the same sweep with the kernels storing into every rank's buffer over NVLink (no all-gather): three sweeps,
        # so that both alternating buffers and the reuse of the first one are exercised, with different volumes
        pg = sharding.PeerGather(2 * 5, 36, 40, dev)
        ok_peer = True
        for it in range(3):
            vol_it = vol * (1.0 + 0.25 * it)
            got = sharding.drr_project_sharded(vol_it, poses, (36, 40), (2.2, 2.2, 2.2), peers=pg)
            ok_peer = ok_peer and bool(torch.equal(got, ops.drr_project(vol_it, poses, (36, 40), (2.2, 2.2, 2.2))))
        pg.close()
        ok_drr = ok_drr and ok_peer
        tp = torch.from_numpy(rs.uniform(-1, 1, (2, 5, 40, 44)).astype(np.float32)).to(dev)
        shape = (25, 20, 28)
        fullv, _ = sharding.backproject_sharded(tp, poses.astype(np.float32), shape, gather=True)
        ok_bp = bool(torch.equal(fullv, ops.backproject(tp, poses.astype(np.float32), shape)))
        img = torch.from_numpy(rs.uniform(-1, 1, (2, 1) + shape).astype(np.float32)).to(dev)
        disp = torch.from_numpy(rs.uniform(-0.1, 0.1, (2, 3) + shape).astype(np.float32)).to(dev)
        slab, zr = sharding.shard_along_z(disp, world, rank)
        wfull = sharding.warp_sharded(img, slab, zr, zero_boundary=True, disp_plus_identity=True, gather=True)
        ok_w = bool(torch.equal(wfull, ops.warp(img, disp, zero_boundary=True, disp_plus_identity=True)))
        q.put((rank, ok_drr, ok_bp, ok_w))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_ops_over_nccl():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(r[1:]) for r in res), res
