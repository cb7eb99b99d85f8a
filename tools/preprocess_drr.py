#!/usr/bin/env python
"""DRR generation step of the reference's tools/preprocessingDRR.py (lines 110-154) on the pipelined B200 path.

    python tools/preprocess_drr.py --preprocessed DATA/preprocessed --task-root DATA/task --out DATA/drr \
        [--phase all|train|debug|val|test] (--scan-range 60 --scan-num 4 | --geo-path geo.csv) [--receptor-size 256 256]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/preprocess_drr.py ...      # cases sharded over the GPUs

Reads `{task_root}/{phase}/data_id.npy` and `{preprocessed}/{id}_{target,source}.npy`, writes `{out}/drr/{id}_{target,source}_proj.npy`
and `{out}/drr/poses.npy` exactly as the reference loop does (same names, shapes, dtypes and values); plotting previews is left to
the reference tool."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from liftreg_b200 import drr_pipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--preprocessed", required=True)
    ap.add_argument("--task-root", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--phase", default="all")
    ap.add_argument("--scan-range", type=float, default=None)
    ap.add_argument("--scan-num", type=int, default=None)
    ap.add_argument("--geo-path", default=None)
    ap.add_argument("--receptor-size", type=int, nargs=2, default=None)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--depth", type=int, default=3, help="cases in flight (pinned staging slots)")
    args = ap.parse_args()
    phases = ["train", "debug", "val", "test"]
    if args.phase in phases:
        phases = [args.phase]
    elif args.phase != "all":
        ap.error("Wrong phase value.")
    drr_folder = os.path.join(args.out, "drr")
    # one process per GPU under torchrun: cases are split round-robin over the ranks (no collective)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device if world == 1 or ":" in args.device else "%s:%s" % (args.device, os.environ.get("LOCAL_RANK", "0"))
    poses = None
    for p in phases:
        ids_path = os.path.join(args.task_root, p, "data_id.npy")
        if not os.path.exists(ids_path):
            print("skipping %s (no %s)" % (p, ids_path))
            continue
        ids = np.load(ids_path)
        t0 = time.perf_counter()
        poses = drr_pipeline.generate_drr_dataset(args.preprocessed, ids, drr_folder, scan_range=args.scan_range,
                                                  scan_num=args.scan_num, geo_path=args.geo_path,
                                                  receptor_size=args.receptor_size, device=device, depth=args.depth,
                                                  shard=(rank, world))
        dt = time.perf_counter() - t0
        print("Processing data in %s ... %d cases in %.2f s (%.1f cases/s)" % (p, len(ids), dt, len(ids) / max(dt, 1e-9)))
    if poses is None:
        print("nothing to do")


if __name__ == "__main__":
    main()
