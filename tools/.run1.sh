#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.txt
tail -15 gpurun_out/r2k_pytest.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
cap() { ncu --cache-control none --clock-control none --metrics $M -k regex:$2 -s 16 -c 32 --csv --log-file gpurun_out/traffic_$1.csv python tools/kbench.py $3 --iters 64 > gpurun_out/traffic_$1.log 2>&1; }
cap backproject_forward_kernel backproject_forward_rows backproject
cap warp_forward_kernel warp_forward warp
cap warp_backward_phi_kernel warp_backward_phi warp_bwd
cap drr_forward_kernel_240x240 drr_forward drr
cap drr_forward_kernel_256x256 drr_forward drr256
cap pca_decode_kernel pca_decode_tma pca
python tools/steady_traffic.py gpurun_out/traffic.json gpurun_out/inst_counts.json backproject_forward_kernel=gpurun_out/traffic_backproject_forward_kernel.csv warp_forward_kernel=gpurun_out/traffic_warp_forward_kernel.csv warp_backward_phi_kernel=gpurun_out/traffic_warp_backward_phi_kernel.csv drr_forward_kernel_240x240=gpurun_out/traffic_drr_forward_kernel_240x240.csv drr_forward_kernel_256x256=gpurun_out/traffic_drr_forward_kernel_256x256.csv pca_decode_kernel=gpurun_out/traffic_pca_decode_kernel.csv
