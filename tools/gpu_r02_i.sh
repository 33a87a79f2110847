#!/bin/bash
# round 2: evidence for profiles/ — ncu captures of the persistent kernels (both modes), launch list of one step, every config
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for p in bf16x3 bf16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_ -s 2 -c 2 -o gpurun_out/r02_persist_$p -f python tools/one_step.py $p 2 > gpurun_out/r02_ncu_$p.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bf16x3.csv python tools/one_step.py bf16x3 1 > /dev/null 2>&1
for c in cfg1 cfg2_linear cfg2_rect cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --cpu-baseline-seconds 10 > gpurun_out/r02_bench_${c}.json 2> gpurun_out/r02_bench_${c}.err
done
timeout 300 python bench.py --config cfg5 --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg5_fp32.json 2> gpurun_out/r02_bench_cfg5_fp32.err
timeout 900 python bench.py --impl reference --config cfg5 --steps 2 --warmup 1 > gpurun_out/r02_bench_cfg5_reference.json 2> gpurun_out/r02_bench_cfg5_reference.err
ls -la gpurun_out | grep r02_ | head -30
