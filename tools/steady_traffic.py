#!/usr/bin/env python
"""Steady-state DRAM traffic and instruction counts per launch from single-pass ncu logs.

    ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
        -k regex:<kernel> -s 16 -c 32 --csv --log-file gpurun_out/traffic_<name>.csv python tools/kbench.py <case> --iters 64
    python tools/steady_traffic.py profiles/traffic.json profiles/inst_counts.json name=log.csv [name=log.csv ...]

Three metrics fit one pass, so nothing is replayed and caches are not flushed (--cache-control none): the launches run
back to back over kbench's 8 rotating buffer sets, and a launch's DRAM counters include the write-back of the dirty
lines earlier launches left in L2 -- the steady state a graph replay sees, unlike a single serialised `--set full`
launch that ends before its own output is written back."""
import csv
import json
import sys


def parse(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    mi, vi, ui, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1, "": 1}
    per = {}
    for r in rows[1:]:
        per.setdefault(r[ki], {})[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    n = len(per)
    mean = lambda m: sum(v.get(m, 0.0) for v in per.values()) / max(n, 1)
    return n, mean("dram__bytes_read.sum"), mean("dram__bytes_write.sum"), mean("smsp__inst_executed.sum")


def main():
    traffic_path, inst_path, items = sys.argv[1], sys.argv[2], sys.argv[3:]
    traffic, inst = {}, {}
    for it in items:
        name, path = it.split("=", 1)
        n, rd, wr, ins = parse(path)
        traffic[name] = int(rd + wr)
        traffic[name + "__read_write"] = [int(rd), int(wr)]
        inst[name] = int(ins)
        print("%-34s launches %3d  read %8.2f MB  write %8.2f MB  inst %.3f M" % (name, n, rd / 1e6, wr / 1e6, ins / 1e6))
    traffic["_how"] = ("mean over 32 back-to-back launches (8 rotating buffer sets), ncu single pass, --cache-control none: "
                       "includes the write-back of earlier launches' dirty L2 lines (steady state)")
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    json.dump(inst, open(inst_path, "w"), indent=1)


if __name__ == "__main__":
    main()
