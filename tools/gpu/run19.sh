cd $GRAFT_REPO_ROOT
timeout 600 python tools/bench_configs.py > gpurun_out/r02_secondary_configs.txt 2>&1; cat gpurun_out/r02_secondary_configs.txt | tail -12
timeout 900 python tools/microbench.py > gpurun_out/r02_microbench.txt 2>&1; tail -5 gpurun_out/r02_microbench.txt
