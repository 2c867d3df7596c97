#!/bin/bash
O=gpurun_out/r2k
mkdir -p $O
for W in 32 48 64 128; do
DZ_ENGINE_WINDOWS=$W timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > $O/bench_w$W.json 2> $O/bench_w$W.err; echo "bench w$W rc=$?" >> $O/summary.txt
done
for E in 16 64; do
DZ_ENGINE_EMB_WINDOWS=$E timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > $O/bench_e$E.json 2> $O/bench_e$E.err; echo "bench e$E rc=$?" >> $O/summary.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2k/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["ms_per_step"],1), d["breakdown"].get("stages_ms"))
    except Exception as e: print(f,"ERR",e)
PY
cat $O/summary.txt
