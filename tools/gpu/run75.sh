cd $GRAFT_REPO_ROOT
timeout 900 python tools/bench_configs.py > gpurun_out/r02_secondary_configs.txt 2>&1; grep -v Warn gpurun_out/r02_secondary_configs.txt | tail -30
timeout 600 python tools/microbench.py > gpurun_out/r02_microbench_c3.txt 2>&1; tail -3 gpurun_out/r02_microbench_c3.txt | cut -c1-300
