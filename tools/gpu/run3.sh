cd $GRAFT_REPO_ROOT
CFG=''
add() { CFG="$CFG${CFG:+;}$1"; }
for d in 2 3 4 5 6 8; do add "{\"depth\": $d, \"env\": {}}"; done
add "{\"depth\": 4, \"env\": {\"PN12_FPS1\": \"8,128,2\"}}"
add "{\"depth\": 6, \"env\": {\"PN12_FPS1\": \"8,128,2\"}}"
timeout 900 python tools/pipeline_sweep.py --steps 96 --configs "$CFG" > gpurun_out/r02_sweep3.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -3 gpurun_out/r02_bench1.err
