cd $GRAFT_REPO_ROOT
python tools/probes/fp1_phase_probe.py 2>&1 | tail -14
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fp_mlp or semseg or graph_replay or blocks_golden or folded" 2>&1 | tail -4
timeout 300 python tools/pipeline_sweep.py --steps 96 --configs '{"depth": 4, "env": {}}' 2>&1 | grep depth
