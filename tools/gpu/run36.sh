cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "cls or msg or heads or partseg or graphed_module" 2>&1 | tail -3
timeout 600 python tools/bench_configs.py 2>&1 | grep "C4"
timeout 300 python tools/probes/c4_timeline.py 2>&1 | grep -v "Warn\|at::native" | tail -28
