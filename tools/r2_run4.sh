RC_TAG=_v4 timeout 200 python tools/r2_ts_check.py time > gpurun_out/ts_time_v4.log 2>&1; echo time rc=$?
grep -h "^time" gpurun_out/ts_time_v4.log | sort -u
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 4 gpurun_out/pytest_gpu.log
