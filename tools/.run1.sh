#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pca or drop_in" > gpurun_out/r2j_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.txt
{
echo "== main";  python tools/kbench.py backproject pca pca_bwd --iters 400
echo "== staged pca"; LIFTREG_B200_PCA_TMA=0 python tools/kbench.py pca_bwd --iters 200
echo "== l1hit"; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/bp_l1hit.so python tools/kbench.py backproject
} > gpurun_out/r2j_kbench.txt 2>&1
tail -3 gpurun_out/r2j_pytest.txt; cat gpurun_out/r2j_kbench.txt
