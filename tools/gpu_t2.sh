#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
LIB=online-neural-cdes_b200/torchcde_b200/libncde_b200.so
cp $LIB /tmp/new.so
for v in old new old new; do
  if [ $v = old ]; then cp tools/micro/libncde_old.so $LIB; else cp /tmp/new.so $LIB; fi
  echo "== $v"; timeout 600 python -m pytest "tests/test_gpu_cfg5_full.py::test_cfg5_full_length_against_oracle" -m gpu -q -s 2>&1 | grep "cfg5 full length\|passed\|failed"
done
cp /tmp/new.so $LIB
