cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_train_launches.csv python tools/bench_train.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_train_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_train_launches.csv 1300 > gpurun_out/r02_train_launches_summary.txt; head -40 gpurun_out/r02_train_launches_summary.txt
