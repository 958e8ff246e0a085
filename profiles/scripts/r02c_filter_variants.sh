#!/bin/bash
# round 2, call 3: tensor filter epilogue variants (dedicated warps per accumulator stage, 32-column TMEM
# loads, survivor emission removed) at bench C2; then the N = 2 bench with the dynamic hash schedule
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in "$@"; do
  HS_LIBRARY=$PWD/hsearch_b200/variants/lib_$v.so timeout 300 python profiles/scripts/filter_time.py 2>&1 | tail -1
done
