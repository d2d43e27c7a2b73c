# full GPU suite + default bench + ncu capture of the evaluation kernel (profiles/r2_eval_topk_ncu.txt)
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_s2d.json 2> gpurun_out/bench_s2d.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2d.json").read().strip().splitlines()[-1])
for k in ("value","roofline","api_sync_free","eval"): print(k, json.dumps(d.get(k))[:900])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_topk -s 2 -c 1 -f -o gpurun_out/r2_topk python tools/prof_topk.py 5 > gpurun_out/r2_topk_ncu.log 2>&1; tail -2 gpurun_out/r2_topk_ncu.log
