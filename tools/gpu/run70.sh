cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "preprocess or empty_scan or chamfer or ragged or eval_iteration" 2>&1 | tail -3
timeout 300 python tools/bench_preprocess.py 2>&1 | tail -3
