#!/bin/bash
# round 2, final state of the persistent kernels: bench line of cfg 5 (full, with the CPU baseline), ncu captures, launch list, traces
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r02m_gputests.txt; cat gpurun_out/r02m_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> gpurun_out/r02m_gputests.txt
for p in bf16x3 bf16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_ -s 2 -c 2 -o gpurun_out/r02m_persist_$p -f python tools/one_step.py $p 2 > gpurun_out/r02m_ncu_$p.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_bf16x3.csv python tools/one_step.py bf16x3 1 > /dev/null 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-baseline-seconds 10 > gpurun_out/r02m_bench_cfg5.json 2> gpurun_out/r02m_bench_cfg5.err
for c in cfg1 cfg2_linear cfg2_rect cfg4; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --cpu-baseline-seconds 5 > gpurun_out/r02m_bench_${c}.json 2> gpurun_out/r02m_bench_${c}.err
done
for m in bf16x3 bf16; do
  NCDE_PS_TRACE_G=0 timeout 200 python tools/ps_trace.py $m 2> gpurun_out/r02m_trace_${m}.txt > /dev/null
  python tools/trace_stats.py gpurun_out/r02m_trace_${m}.txt > gpurun_out/r02m_trace_${m}_summary.txt 2>&1
  python tools/trace_all.py gpurun_out/r02m_trace_${m}.txt >> gpurun_out/r02m_trace_${m}_summary.txt 2>&1
done
tail -c 400 gpurun_out/r02m_bench_cfg5.json
