cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -5 gpurun_out/r02_bench1.err
