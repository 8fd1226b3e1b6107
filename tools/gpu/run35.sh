cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "fps" 2>&1 | tail -2
timeout 900 python tools/microbench.py > gpurun_out/r02_microbench.txt 2>&1; tail -2 gpurun_out/r02_microbench.txt | cut -c1-150
