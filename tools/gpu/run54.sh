cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "three_nn or nn_blocks or interpol or fp or semseg or partseg or golden" 2>&1 | tail -3
for bp in 8 16 32; do PN12_NN_BLOCK=$bp python tools/probes/nn1_ab.py 2>&1 | grep variant | sed "s/^/BP=$bp /"; done
for bp in 8 16 32; do PN12_NN_BLOCK=$bp timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}}' 2>&1 | grep "depth\": 10" | sed "s/^/BP=$bp /"; done
