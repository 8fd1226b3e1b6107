cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "fps" 2>&1 | tail -3
timeout 900 python tools/microbench.py 2>&1 | grep '"N": 65536\|"N": 120000' | cut -c1-140
