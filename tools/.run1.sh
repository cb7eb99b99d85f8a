#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.txt
{
echo "== main";  python tools/kbench.py warp warp_bwd backproject drr
echo "== main exact drr"; LIFTREG_B200_NUMERICS=exact python tools/kbench.py drr
for v in bp_isub32 bp_u8 bp_u8isub32 bp_u6; do echo "== $v"; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py backproject; done
for v in drr5 drr6 drr_p4; do echo "== $v"; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py drr; done
} > gpurun_out/r2h_kbench.txt 2>&1
tail -3 gpurun_out/r2h_pytest.txt; cat gpurun_out/r2h_kbench.txt
