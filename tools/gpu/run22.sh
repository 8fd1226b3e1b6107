cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu.log
