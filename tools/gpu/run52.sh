cd $GRAFT_REPO_ROOT
for v in 0 1 0 1; do PN12_NN_VARIANT=$v python tools/probes/nn1_ab.py 2>&1 | grep variant; done
for v in 0 1 0 1; do PN12_NN_VARIANT=$v timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}}' 2>&1 | grep "depth\": 10"; done
