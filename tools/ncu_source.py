#!/usr/bin/env python
"""Per-opcode / per-region digest of the source page of one ncu report (needs --import-source on, -lineinfo).

    python tools/ncu_source.py rep.ncu-rep [--top N] [--regions]

Prints: warp instructions executed and stall samples by opcode class, the hottest instructions by stall samples, and
(--regions) contiguous address ranges with a similar execution count (set-up / loop bodies)."""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    col = {n: hdr.index(n) for n in hdr}

    def num(r, name):
        if name not in col:            # kernels without shared memory have no shared-memory columns
            return 0.0
        v = r[col[name]]
        try:
            return float(v)
        except ValueError:
            return 0.0

    tot = sum(num(r, "Instructions Executed") for r in data)
    tot_st = sum(num(r, "Warp Stall Sampling (All Samples)") for r in data)
    print("kernel:", rows[0][1][:100])
    print("warp instructions executed: %.3f M; stall samples: %d; static instructions: %d" % (tot / 1e6, tot_st, len(data)))
    by = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    for r in data:
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.rstrip(";").split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "STG", "LDS", "STS")) and "." in op else "")
        b = by[op]
        b[0] += num(r, "Instructions Executed")
        b[1] += num(r, "Warp Stall Sampling (All Samples)")
        b[2] += num(r, "L1 Tag Requests Global")
        b[3] += num(r, "L1 Wavefronts Shared")
    print("\n%-14s %10s %7s %9s %7s %12s %12s" % ("opcode", "exec (M)", "%", "stalls", "%", "L1 tag req", "smem wavefr"))
    for op, b in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-14s %10.3f %6.1f%% %9d %6.1f%% %12d %12d" % (op, b[0] / 1e6, 100 * b[0] / tot, b[1], 100 * b[1] / max(tot_st, 1), b[2], b[3]))
    print("\nhottest instructions by stall samples:")
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for r in sorted(data, key=lambda r: -num(r, "Warp Stall Sampling (All Samples)"))[:top]:
        reasons = sorted(((num(r, n), n[6:]) for n in stall_cols), reverse=True)[:2]
        print("  %6d  exec %8d  %-70s %s" % (num(r, "Warp Stall Sampling (All Samples)"), num(r, "Instructions Executed"),
                                              r[col["Source"]].strip()[:70], ", ".join("%s=%d" % (n, v) for v, n in reasons if v)))
    if "--regions" in sys.argv:
        print("\nregions (contiguous instructions with similar execution counts):")
        cur = None
        for k, r in enumerate(data):
            c = num(r, "Instructions Executed")
            s = num(r, "Warp Stall Sampling (All Samples)")
            if cur and abs(c - cur[2]) <= 0.15 * max(c, cur[2], 1):
                cur[1] = k; cur[3] += c; cur[4] += s; cur[5] += 1
            else:
                if cur and cur[3] > 0.004 * tot:
                    print("  instr %4d-%4d  n=%3d  exec/instr ~%9d  sum %.3f M (%.1f%%)  stalls %d" % (cur[0], cur[1], cur[5], cur[2], cur[3] / 1e6, 100 * cur[3] / tot, cur[4]))
                cur = [k, k, c, c, s, 1]
        if cur and cur[3] > 0.004 * tot:
            print("  instr %4d-%4d  n=%3d  exec/instr ~%9d  sum %.3f M (%.1f%%)  stalls %d" % (cur[0], cur[1], cur[5], cur[2], cur[3] / 1e6, 100 * cur[3] / tot, cur[4]))


if __name__ == "__main__":
    main()
