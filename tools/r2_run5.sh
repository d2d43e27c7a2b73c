timeout 300 python tools/prof_api.py > gpurun_out/prof_api.log 2>&1; echo api rc=$?; tail -n 14 gpurun_out/prof_api.log
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-eval --no-area --no-shared > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2b.json'))
for k in ('value','ms_per_step','ts_kernel','api_device','hybrid','kcliff','with_dtext'): print(k, json.dumps(d.get(k))[:600])
PY
