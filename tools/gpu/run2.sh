cd $GRAFT_REPO_ROOT
CFG=''
add() { CFG="$CFG${CFG:+;}$1"; }
for d in 3 4; do
 for c in 8 16 24 32 48; do
  add "{\"depth\": $d, \"env\": {\"PN12_FPS1\": \"4,256,2\", \"PN12_STREAM_BALL_CTAS\": \"$c\"}}"
 done
 add "{\"depth\": $d, \"env\": {\"PN12_FPS1\": \"4,256,2\", \"PN12_STREAM_BALL_CTAS\": \"16\", \"PN12_STREAM_BALL_SHARE\": \"1\"}}"
 add "{\"depth\": $d, \"env\": {\"PN12_FPS1\": \"4,256,2\", \"PN12_STREAM_BALL_CTAS\": \"32\", \"PN12_STREAM_BALL_SHARE\": \"1\"}}"
 add "{\"depth\": $d, \"env\": {\"PN12_FPS1\": \"8,256,2\", \"PN12_STREAM_BALL_CTAS\": \"24\"}}"
 add "{\"depth\": $d, \"env\": {\"PN12_STREAM_BALL_CTAS\": \"16\"}}"
 add "{\"depth\": $d, \"env\": {\"PN12_FPS1\": \"4,256,3\", \"PN12_STREAM_BALL_CTAS\": \"24\"}}"
done
TIMELINE_ENV='{"PN12_FPS1": "4,256,2", "PN12_STREAM_BALL_CTAS": "24"}' timeout 900 python tools/pipeline_sweep.py --steps 60 --configs "$CFG" --timeline 3 > gpurun_out/r02_sweep2.txt 2>&1
