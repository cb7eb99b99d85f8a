#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LIFTREG_B200_BP_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backproject or cfg3 or slab or host_entry or drop_in or 320" > gpurun_out/r2b_pytest_tma.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_tma.txt
{
echo "== fast";  python tools/kbench.py backproject
echo "== fast + TMA out"; LIFTREG_B200_BP_TMA=1 python tools/kbench.py backproject
echo "== fast + TMA batch 8"; LIFTREG_B200_BP_TMA=1 python tools/kbench.py backproject --batch 8 --iters 400
} > gpurun_out/r2b_kbench.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:backproject_forward_rows -s 3 -c 1 -o gpurun_out/prof_r2b_bp python tools/kbench.py backproject --iters 16 > gpurun_out/ncu_r2b_bp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_forward -s 3 -c 1 -o gpurun_out/prof_r2b_warp python tools/kbench.py warp --iters 16 > gpurun_out/ncu_r2b_warp.log 2>&1
LIFTREG_B200_BP_TMA=1 ncu --set full --clock-control none --import-source on -k regex:backproject_forward_rows -s 3 -c 1 -o gpurun_out/prof_r2b_bp_tma python tools/kbench.py backproject --iters 16 > gpurun_out/ncu_r2b_bp_tma.log 2>&1
tail -3 gpurun_out/r2b_pytest_tma.txt; cat gpurun_out/r2b_kbench.txt
