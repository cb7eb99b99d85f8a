ncu --set full --clock-control none --import-source on -k regex:backproject_forward -c 1 -o gpurun_out/prof_bp_i python tools/kbench.py backproject --iters 8 > gpurun_out/ncu_bp_i.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_forward -c 1 -o gpurun_out/prof_warp_e python tools/kbench.py warp --iters 8 > gpurun_out/ncu_warp_e.log 2>&1
