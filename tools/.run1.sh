set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kbench.py backproject pca_bwd pca 2>&1 | tail -4
for v in isub8 isub32; do LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py backproject 2>&1 | tail -1; done
