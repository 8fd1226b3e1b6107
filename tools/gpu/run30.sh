cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "sa_mlp or semseg or cls or blocks or msg or heads or pipelined or small_levels or mlp_rows or resident or n_sliced or golden" 2>&1 | tail -3
python tools/probes/res_timeline.py sa1 2>&1 | grep -v Warn | head -1
python tools/probes/res_timeline.py sa2 2>&1 | grep -v Warn | head -1
timeout 600 python tools/bench_configs.py 2>&1 | grep "C4" | head -2
timeout 300 python tools/pipeline_sweep.py --steps 96 --configs '{"depth": 6, "env": {}}' 2>&1 | grep depth
