#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in bp_head bp_chunkflag bp_chunkflag_neither bp_chunkflag_noRC; do echo "== $v"; LIFTREG_B200_LIB=$PWD/liftreg_b200/_lib/variants/$v.so python tools/kbench.py backproject; done
