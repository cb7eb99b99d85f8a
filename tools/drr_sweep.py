#!/usr/bin/env python
"""BASELINE.json configs[3]: DRR generation sweep -- 512^3 CT, 64 views over 60 deg, 512^2 detector, view-sharded across
the GPUs of one box (liftreg_b200.sharding.drr_project_sharded); the detector images reach every rank either through P2P
stores from the DRR kernel (sharding.PeerGather), through an NCCL all-gather, or stay sharded.

    python tools/drr_sweep.py                                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/drr_sweep.py

Prints one JSON line (rank 0): nominal ray-samples/s for the whole job, device time = max over ranks (CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from liftreg_b200 import sharding, synthetic  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, P, det = 512, 64, (512, 512)
    z = np.arange(n, dtype=np.float32)
    vol = (0.1 + 0.05 * np.sin(z / 13.0)[:, None, None] * np.cos(z / 17.0)[None, :, None]
           + 0.04 * np.sin(z / 11.0)[None, None, :]).astype(np.float32)
    tv = torch.from_numpy(vol[None]).to(dev)                       # replicated volume (537 MB)
    poses = synthetic.wrapper_poses(60.0, P, n)
    reps = 5
    pg = sharding.PeerGather(P, det[0], det[1], dev)
    for exchange in ("p2p_stores", "nccl_all_gather", "none"):
        gather = exchange != "none"
        kw = {"peers": pg} if exchange == "p2p_stores" else {"gather": gather}
        for _ in range(2):
            out = sharding.drr_project_sharded(tv, poses, det, (1.0, 1.0, 1.0), **kw)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = sharding.drr_project_sharded(tv, poses, det, (1.0, 1.0, 1.0), **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rank == 0:
            nominal = P * det[0] * det[1] * n
            full = out if gather else out[0]          # (gather=False: this rank's images and its view list)
            print(json.dumps({"workload": "cfg4: DRR sweep 512^3, 64 views / 60 deg, 512^2 detector", "n_gpus": world,
                              "exchange": exchange, "ms_per_sweep": ms, "nominal_ray_samples": nominal,
                              "samples_per_s": nominal / ms * 1e3, "checksum": float(full.double().sum().item())}))
    pg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
