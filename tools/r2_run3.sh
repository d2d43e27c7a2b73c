export RANGECLIP_B200_LIB=$PWD/rangeclip_b200/librangeclip_b200_bringup.so
timeout 200 python tools/r2_ts_check.py parity > gpurun_out/ts_parity.log 2>&1; echo parity rc=$?
grep -h "^parity" gpurun_out/ts_parity.log | cut -c1-330; tail -n 3 gpurun_out/ts_parity.log | cut -c1-300
RC_TAG=_v3b timeout 200 python tools/r2_ts_check.py time > gpurun_out/ts_time_v3b.log 2>&1; echo time rc=$?
grep -h "^time" gpurun_out/ts_time_v3b.log; tail -n 2 gpurun_out/ts_time_v3b.log | cut -c1-300
unset RANGECLIP_B200_LIB
RC_TAG=_v3 timeout 200 python tools/r2_ts_check.py time > gpurun_out/ts_time_v3.log 2>&1; echo time rc=$?
grep -h "^time" gpurun_out/ts_time_v3.log; tail -n 2 gpurun_out/ts_time_v3.log | cut -c1-300
