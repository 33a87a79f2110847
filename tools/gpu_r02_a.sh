#!/bin/bash
# round 2, call A: regression (all GPU tests incl. the new full-length oracle tests), widened-mode throughput, every config
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1
timeout 400 python tools/bench_modes.py --steps 3 > gpurun_out/r02a_bench_modes.jsonl 2> gpurun_out/r02a_bench_modes.err
for c in cfg1 cfg2_linear cfg2_rect cfg3 cfg4; do
  for p in bf16 fp32; do
    timeout 300 python bench.py --config $c --precision $p --steps 5 --warmup 3 --cpu-baseline-seconds 8 > gpurun_out/r02a_bench_${c}_${p}.json 2> gpurun_out/r02a_bench_${c}_${p}.err
  done
done
timeout 300 python bench.py --config cfg5 --precision bf16 --steps 10 --warmup 3 > gpurun_out/r02a_bench_cfg5_bf16.json 2> gpurun_out/r02a_bench_cfg5_bf16.err
timeout 300 python bench.py --config cfg5 --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_cfg5_fp32.json 2> gpurun_out/r02a_bench_cfg5_fp32.err
timeout 600 python bench.py --impl reference --config cfg5 --steps 2 --warmup 1 > gpurun_out/r02a_bench_cfg5_reference.json 2> gpurun_out/r02a_bench_cfg5_reference.err
timeout 120 python bench.py --check > gpurun_out/r02a_check1.json 2>&1
tail -3 gpurun_out/r02a_tests.log
