cd $GRAFT_REPO_ROOT
CFG=''
add() { CFG="$CFG${CFG:+;}$1"; }
add "{\"depth\": 6, \"env\": {}}"
add "{\"depth\": 6, \"env\": {\"PN12_NN1_BACKGROUND\": \"0\"}}"
add "{\"depth\": 6, \"env\": {\"PN12_RESERVE_L2\": \"0\"}}"
add "{\"depth\": 6, \"env\": {\"PN12_NN1_BACKGROUND\": \"0\", \"PN12_RESERVE_L2\": \"0\"}}"
add "{\"depth\": 6, \"env\": {}}"
timeout 900 python tools/pipeline_sweep.py --steps 96 --configs "$CFG" 2>&1 | grep depth
