#!/bin/bash
O=gpurun_out/r2e
mkdir -p $O
timeout 600 python scripts/dbg_clustering.py > $O/dbg_clustering.log 2>&1; echo "dbg rc=$?" >> $O/summary.txt
cat $O/dbg_clustering.log
timeout 1200 python -m pytest tests/test_emb_gpu.py tests/test_gemm_gpu.py tests/test_post_gpu.py tests/test_seg_gpu.py tests/test_bench_config_gpu.py -q --timeout 900 > $O/t_kernels.log 2>&1; echo "kernel tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED" $O/t_kernels.log | tail -20
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e/bench.json"))
print(d["value"], d["ms_per_step"], d["breakdown"].get("stages_ms"))
p=json.load(open("gpurun_out/r2e/prof.json"))
print({x["name"]: round(x["ms"],3) for x in p["seg"] if x["name"] in ("conv0","conv1","conv2","conv3","C0_dwconv")})
print({x["name"]: round(x["ms"],3) for x in p["emb"] if x["name"].startswith("l3b1") or x["name"].startswith("l3b0")})
PY
cat $O/summary.txt
