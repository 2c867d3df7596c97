#!/bin/bash
O=gpurun_out/r2g
mkdir -p $O
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_seg_gpu.py tests/test_bench_config_gpu.py tests/test_emb_gpu.py tests/test_pipeline_gpu.py -q --timeout 900 > $O/t_kernels.log 2>&1; echo "kernel tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED" $O/t_kernels.log | tail -20
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
DZ_GEMM_NO_DEEP=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --profile-out $O/prof_nodeep.json > $O/bench_nodeep.json 2> $O/bench_nodeep.err; echo "bench nodeep rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --precision bf16x3 --no-cpu-baseline --no-sub-records > $O/bench_bf16x3.json 2> $O/bench_bf16x3.err; echo "bench bf16x3 rc=$?" >> $O/summary.txt
python - <<'PY'
import json
for f in ("bench","bench_nodeep","bench_bf16x3"):
    try:
        d=json.load(open(f"gpurun_out/r2g/{f}.json"))
        print(f, round(d["value"],1), round(d["ms_per_step"],1), d["breakdown"].get("stages_ms"), {k:v.get("tflops") for k,v in d["roofline"]["groups"].items() if k.startswith("wavlm") or k.startswith("cnn")})
    except Exception as e: print(f, "ERR", e)
PY
cat $O/summary.txt
