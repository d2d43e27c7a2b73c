timeout 200 python tools/r2_ts_check.py parity > gpurun_out/ts_parity.log 2>&1; echo parity rc=$?
for spb in 2 4 8; do RC_TAG=_spb$spb RANGECLIP_B200_LIB=$PWD/rangeclip_b200/librangeclip_b200_bringup.so RANGECLIP_B200_TS_SPB=$spb timeout 200 python tools/r2_ts_check.py time > gpurun_out/ts_time_spb$spb.log 2>&1; echo spb$spb rc=$?; done
grep -h "^time" gpurun_out/ts_time_spb*.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
tail -n 15 gpurun_out/pytest_gpu.log
