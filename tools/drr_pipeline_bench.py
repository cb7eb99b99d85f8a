#!/usr/bin/env python
"""Row f3 measurement: the dataset DRR loop (tools/preprocessingDRR.py:123-154) as the pipeline of
liftreg_b200/drr_pipeline.py vs the same loop written like the reference (two synchronous mirror calls + np.save per case).

    python tools/drr_pipeline_bench.py [--cases 24] [--dir /dev/shm/lr_drr_bench]

Synthetic 160^3 HU volumes (float32 .npy, like the reference's preprocessed data), 4 views over 60 degrees, 240^2 detector."""
import argparse
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from liftreg_b200 import drr_pipeline, sdct_projection_utils as sdct, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=24)
    ap.add_argument("--dir", default="/dev/shm/lr_drr_bench")
    args = ap.parse_args()
    pre, out_a, out_b = (os.path.join(args.dir, d) for d in ("pre", "pipeline", "serial"))
    shutil.rmtree(args.dir, ignore_errors=True)
    os.makedirs(pre)
    base = synthetic.ct_phantom((160, 160, 160), sigma=1.0, nodules=8)
    rs = np.random.RandomState(0)
    ids = ["c%03d" % i for i in range(args.cases)]
    for i in ids:
        for kind in ("target", "source"):
            np.save(os.path.join(pre, "%s_%s.npy" % (i, kind)), (base + rs.normal(0, 5, base.shape)).astype(np.float32))

    drr_pipeline.generate_drr_dataset(pre, ids[:2], out_a, scan_range=60.0, scan_num=4)            # warm-up (module load)
    t0 = time.perf_counter()
    drr_pipeline.generate_drr_dataset(pre, ids, out_a, scan_range=60.0, scan_num=4)
    t_pipe = time.perf_counter() - t0

    os.makedirs(out_b)
    t0 = time.perf_counter()
    for i in ids:                                                # the reference loop, calling the mirror functions
        target = np.flip(np.load(os.path.join(pre, "%s_target.npy" % i)), axis=1)
        source = np.flip(np.load(os.path.join(pre, "%s_source.npy" % i)), axis=1)
        sp, poses = sdct.calculate_projection_wraper(sdct.calc_relative_atten_coef(source), 60.0, 4, (2.2, 2.2, 2.2))
        tp, _ = sdct.calculate_projection_wraper(sdct.calc_relative_atten_coef(target), 60.0, 4, (2.2, 2.2, 2.2))
        np.save(os.path.join(out_b, "%s_target_proj.npy" % i), tp)
        np.save(os.path.join(out_b, "%s_source_proj.npy" % i), sp)
    t_serial = time.perf_counter() - t0
    same = all(np.array_equal(np.load(os.path.join(out_a, f)), np.load(os.path.join(out_b, f))) for f in os.listdir(out_b))
    print(json.dumps({"cases": args.cases, "pipeline_s": t_pipe, "pipeline_cases_per_s": args.cases / t_pipe,
                      "serial_mirror_s": t_serial, "serial_cases_per_s": args.cases / t_serial,
                      "speedup": t_serial / t_pipe, "identical_files": bool(same),
                      "storage": args.dir, "volume": "160^3 float32 x2 per case, 4 views, 240^2"}))
    shutil.rmtree(args.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
