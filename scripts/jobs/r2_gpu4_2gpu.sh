#!/bin/bash
# 2-GPU job: the window-sharded single-recording bench (strong scaling headline) + RTTM equality check, NCCL
O=gpurun_out/r2d
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench 2gpu rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d/bench_2gpu.json"))
print(d["value"], d["ms_per_step"], d["scaling"], d["e2e"], d["config"]["sharded_rttm_equals_unsharded"], d["config"]["replicas"], d["config"]["clusters_found"])
PY
tail -5 $O/bench_2gpu.err
cat $O/summary.txt
