python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kbench.py backproject warp warp_bwd 2>&1 | tail -3
for v in A B C D E; do echo $v; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py backproject warp 2>&1 | tail -2; done
