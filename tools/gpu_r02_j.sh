#!/bin/bash
# round 2, second half: evidence for profiles/ after the backward-pass pipelining — full GPU test suite, ncu captures of the persistent
# kernels (both modes), launch list of one step, every config, hand-off traces, MMA issue-rate micro-benchmark
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02j_gputests.txt; cat gpurun_out/r02j_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/r02j_smoke.txt; cat gpurun_out/r02j_smoke.txt
for p in bf16x3 bf16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_ -s 2 -c 2 -o gpurun_out/r02j_persist_$p -f python tools/one_step.py $p 2 > gpurun_out/r02j_ncu_$p.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches_bf16x3.csv python tools/one_step.py bf16x3 1 > /dev/null 2>&1
for c in cfg1 cfg2_linear cfg2_rect cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --cpu-baseline-seconds 10 > gpurun_out/r02j_bench_${c}.json 2> gpurun_out/r02j_bench_${c}.err
done
timeout 300 python bench.py --config cfg5 --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_cfg5_fp32.json 2> gpurun_out/r02j_bench_cfg5_fp32.err
for m in bf16x3 bf16; do
  NCDE_PS_TRACE_G=0 timeout 200 python tools/ps_trace.py $m 2> gpurun_out/r02j_trace_${m}.txt > /dev/null
  python tools/trace_stats.py gpurun_out/r02j_trace_${m}.txt > gpurun_out/r02j_trace_${m}_summary.txt 2>&1
  python tools/trace_all.py gpurun_out/r02j_trace_${m}.txt >> gpurun_out/r02j_trace_${m}_summary.txt 2>&1
done
timeout 60 tools/micro/mma_rate > gpurun_out/r02j_mma_rate.txt 2>&1
ls -la gpurun_out | grep r02j_ | head -40
