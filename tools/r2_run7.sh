timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 3 gpurun_out/pytest_gpu.log
for v in shared2x2 full; do timeout 600 python bench.py --workload full_step --variant $v --steps 6 --warmup 3 > gpurun_out/full1_$v.json 2> gpurun_out/full1_$v.err; echo full $v rc=$?; python -c "
import json; d=json.load(open('gpurun_out/full1_$v.json')); print({k:d[k] for k in ('value','ms_per_step','backbone_only_ms','loss_path_share')})"; done
timeout 600 python bench.py --workload full_step --variant eager --batch 16 --steps 4 --warmup 2 > gpurun_out/full1_eager.json 2> gpurun_out/full1_eager.err; echo eager rc=$?; tail -c 600 gpurun_out/full1_eager.err; python -c "
import json; d=json.load(open('gpurun_out/full1_eager.json')); print({k:d[k] for k in ('value','ms_per_step','backbone_only_ms','loss_path_share','config')})"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-eval --no-area --no-shared > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
for k in ('value','ms_per_step','api_device','hybrid','kcliff'): print(k, json.dumps(d.get(k))[:500])
PY
