cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "ball or golden or semseg or stream or pipelined" 2>&1 | tail -2
python tools/probes/bq1_probe.py 2>&1 | grep "ball query" | head -1
timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {}}' 2>&1 | grep "depth"
timeout 600 python tools/microbench.py 2>&1 | grep -i "ball" | head -20
