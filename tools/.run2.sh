python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kbench.py backproject 2>&1 | tail -1
for v in g5 g4 g3 isub8; do echo $v; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py backproject 2>&1 | tail -1; done
ncu --set full --clock-control none --import-source on -k regex:backproject_forward -c 1 -o gpurun_out/prof_bp_h python tools/kbench.py backproject --iters 8 > gpurun_out/ncu_bp_h.log 2>&1
