#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.txt
tail -5 gpurun_out/r2y_pytest.txt
python tools/kbench.py warp warp_bwd
python tools/kbench.py warp warp_bwd --batch 8 --iters 400
for rep in 1 2; do
echo "== main"; python tools/kbench.py backproject; python tools/kbench.py backproject --batch 8 --iters 400
for v in bp_noRC bp_noRC_noMagic; do echo "== $v"; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py backproject; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py backproject --batch 8 --iters 400; done
done
