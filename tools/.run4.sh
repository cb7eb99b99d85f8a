python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kbench.py backproject warp 2>&1 | tail -2
for t in "0,100" "0,0" "40,40" "60,25" "75,15" "100,0" "30,30" "50,50"; do echo "taper $t"; LIFTREG_B200_WARP_TAPER=$t LIFTREG_B200_BP_TAPER=$t python tools/kbench.py backproject warp 2>&1 | tail -2; done
