#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02e_tests.log
timeout 300 python -m pytest tests/test_gpu_cfg5_full.py -q -s 2>&1 | grep -E "cfg5 full|training|passed|failed|Error|error" > gpurun_out/r02e_cfg5.log
for p in bf16 bf16x3; do
  timeout 300 python bench.py --config cfg5 --precision $p --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_cfg5_$p.json 2> gpurun_out/r02e_bench_cfg5_$p.err
done
for c in cfg1 cfg2_linear cfg2_rect cfg4; do
  timeout 300 python bench.py --config $c --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_${c}_bf16.json 2> gpurun_out/r02e_bench_${c}_bf16.err
done
tail -4 gpurun_out/r02e_tests.log; cat gpurun_out/r02e_cfg5.log
