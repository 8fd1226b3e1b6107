cd $GRAFT_REPO_ROOT
CFG=''
add() { CFG="$CFG${CFG:+;}$1"; }
add "{\"depth\": 4, \"env\": {}}"
for w in fps1 bq1 nn1 fp1 sa1 sa2 "fps1,fp1" "fps1,bq1,nn1" "sa1,sa2,fp1" "fps1,bq1,nn1,fp1,sa1,sa2"; do add "{\"depth\": 4, \"env\": {\"PN12_WHATIF\": \"$w\"}}"; done
timeout 900 python tools/pipeline_sweep.py --steps 96 --configs "$CFG" 2>&1 | grep depth
