#!/bin/bash
O=gpurun_out/r2q
mkdir -p $O
timeout 900 python -m pytest tests/test_emb_gpu.py tests/test_bench_config_gpu.py tests/test_glue_golden_gpu.py -q --timeout 600 > $O/t_emb.log 2>&1; echo "emb tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED" $O/t_emb.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2q/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],1), d["breakdown"].get("stages_ms"), d["breakdown"]["emb:other"])
p=json.load(open("gpurun_out/r2q/prof.json"))
print({x["name"]: round(x["ms"],3) for x in p["emb"] if x["name"] in ("fbank","fbank_mean","conv1","stats_pool","seg_1")})
PY
cat $O/summary.txt
