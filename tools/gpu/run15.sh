cd $GRAFT_REPO_ROOT
CFG=''
add() { CFG="$CFG${CFG:+;}$1"; }
add "{\"depth\": 4, \"env\": {}}"
for c in "0,256,0" "0,512,0" "0,0,0"; do for d in 4 6; do add "{\"depth\": $d, \"env\": {\"PN12_FPS1_SORTED\": \"$c\"}}"; done; done
timeout 900 python tools/pipeline_sweep.py --steps 96 --configs "$CFG" 2>&1 | grep depth
