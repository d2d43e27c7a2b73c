timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_full_step.py shared2x2 > gpurun_out/prof_fs_shared.log 2>&1; echo rc=$?
timeout 300 python tools/prof_full_step.py full > gpurun_out/prof_fs_full.log 2>&1; echo rc=$?
timeout 300 python tools/prof_api.py > gpurun_out/prof_api.log 2>&1; echo api rc=$?; tail -n 14 gpurun_out/prof_api.log
