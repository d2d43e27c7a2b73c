# timing build: per-role wait breakdown + event timeline of the pair kernel; then ncu captures for profiles/r2_*
python tools/timing_pair.py 64 > gpurun_out/s2_timing_pair.txt 2>&1; echo timing rc=$?
python tools/timeline_pair.py > gpurun_out/s2_timeline_pair.txt 2>&1; echo timeline rc=$?
SS=1 PROF_REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_umma_pair -s 1 -c 1 -f -o gpurun_out/r2_pair python tools/prof_ts.py > gpurun_out/r2_pair_ncu.log 2>&1; echo ncu pair rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-eager > gpurun_out/r2_launches_bench.log 2>&1; echo ncu launches rc=$?
