#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.txt
tail -6 gpurun_out/r2m_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/kbench.py drr drr256 warp warp_bwd backproject pca_bwd
LIFTREG_B200_NUMERICS=exact python tools/kbench.py drr
