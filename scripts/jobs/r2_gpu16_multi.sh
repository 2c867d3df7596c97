#!/bin/bash
# N-GPU job: window-sharded single-recording headline at N ranks, then the recording-list workload
N=${1:-8}
O=gpurun_out/r2n$N
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload many --recordings 18 --minutes 10 --steps 1 > $O/bench_many.json 2> $O/bench_many.err; echo "bench many rc=$?" >> $O/summary.txt
python - <<PY
import json
for f in ("bench","bench_many"):
    try:
        d=json.load(open("$O/%s.json" % f))
        c=d["config"]
        print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],1), d["scaling"], c.get("sharded_rttm_equals_unsharded"), c.get("root_window_share"), c.get("replicas"), c.get("annotations_returned"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $O/bench.err; tail -3 $O/bench_many.err
cat $O/summary.txt
