cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "graphed_module or seg_c1" 2>&1 | tail -3
timeout 600 python tools/bench_configs.py 2>&1 | grep "C4\|C1"
