cd $GRAFT_REPO_ROOT
CFG=''
add() { CFG="$CFG${CFG:+;}$1"; }
for d in 4 5 6; do add "{\"depth\": $d, \"env\": {}}"; add "{\"depth\": $d, \"env\": {\"PN12_FPS1\": \"4,256,3\"}}"; done
timeout 900 python tools/pipeline_sweep.py --steps 96 --configs "$CFG" 2>&1 | grep depth
