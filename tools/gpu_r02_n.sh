#!/bin/bash
# 4-GPU weak scaling + NCCL gradient check with the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --check > gpurun_out/r02n_check$N.json 2> gpurun_out/r02n_check$N.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-alt-mode > gpurun_out/r02n_bench_${N}gpu.json 2> gpurun_out/r02n_bench_${N}gpu.err
tail -1 gpurun_out/r02n_check$N.json | cut -c1-300; tail -c 300 gpurun_out/r02n_check$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02n_bench_%sgpu.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("n", d["n_gpus"], "ms", round(d["ms_per_step"], 2), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["scaling"])
PY
