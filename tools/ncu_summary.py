#!/usr/bin/env python
"""Summarise ncu reports (one or more .ncu-rep) into a text table + traffic.json (DRAM bytes per launch).

    python tools/ncu_summary.py profiles/r01_ncu_full_summary.txt profiles/traffic.json rep1.ncu-rep [rep2 ...]
"""
import csv
import json
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
SCALE = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9}


def main():
    out_txt, out_json, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines, traffic, seen = [], {}, set()
    for rep in reps:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ki = hdr.index('Kernel Name')
        for r in rows[2:]:
            name = re.sub(r'[<(].*', '', r[ki].replace('void ', '')).strip().split('::')[-1]
            if name in seen:
                continue
            seen.add(name)
            lines.append('===== ' + r[ki][:120] + '   [' + rep.split('/')[-1] + ']')
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append('  %-88s %16s %s' % (w, r[i], units[i]))

            def val(m):
                i = hdr.index(m)
                return float(r[i]) * SCALE[units[i]]
            traffic[name] = int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum'))
    head = ("ncu --set full --clock-control none --import-source on, one launch per kernel (cold caches, serialised): compare\n"
            "shares and ratios, not absolute times (CUDA-event times are in the bench JSON).  B200, cfg 2 / cfg 1 sizes via\n"
            "tools/kbench.py.  traffic.json = dram__bytes_read.sum + dram__bytes_write.sum per launch.\n\n")
    open(out_txt, 'w').write(head + '\n'.join(lines) + '\n')
    json.dump(traffic, open(out_json, 'w'), indent=1)
    print('\n'.join(lines))
    print(traffic)


if __name__ == '__main__':
    main()
